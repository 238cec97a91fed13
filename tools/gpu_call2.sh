#!/bin/bash
# round 2, GPU call 2: GPU tests after the fixes (graphs, features, frame loop), bench with sub_results
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -12 gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"
tail -c 800 gpurun_out/r2b_bench.err
timeout 300 python tools/phase_times.py C2 1 > gpurun_out/r2b_phase_c2_1.txt 2>&1
echo done
