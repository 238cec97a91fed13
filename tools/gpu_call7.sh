#!/bin/bash
# round 2, GPU call 7: TMA-fed map accumulation: tests, phase timings, bench
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/r2f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -8 gpurun_out/r2f_pytest.log
timeout 300 python tools/phase_times.py C2 4 > gpurun_out/r2f_phase_c2.txt 2>&1
timeout 300 python tools/phase_times.py C3 4 > gpurun_out/r2f_phase_c3.txt 2>&1
grep -h "maps accumulate\|gaussian" gpurun_out/r2f_phase_c2.txt gpurun_out/r2f_phase_c3.txt
timeout 300 python tools/layer_times.py coco 16 736 984 > gpurun_out/r2f_layers_coco_16.txt 2>&1
timeout 300 python tools/layer_times.py hand 32 736 736 > gpurun_out/r2f_layers_hand_32.txt 2>&1
head -4 gpurun_out/r2f_layers_coco_16.txt gpurun_out/r2f_layers_hand_32.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2f_smoke.txt 2>&1
tail -2 gpurun_out/r2f_smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"
tail -c 300 gpurun_out/r2f_bench.err
echo done
