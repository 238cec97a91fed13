#!/bin/bash
# round 2, GPU call 4: GPU tests, bench with graphs, single-frame phases, then the ncu captures
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
timeout 300 python tools/phase_times.py C2 1 > gpurun_out/r2c_phase_c2_1.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
echo "bench rc=$?"
bash tools/gpu_ncu.sh
