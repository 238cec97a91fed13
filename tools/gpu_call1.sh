#!/bin/bash
# round 2, GPU call 1: all GPU tests, the bench line with sub_results, phase and layer timings
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench.err
timeout 300 python tools/phase_times.py C2 4 > gpurun_out/r2a_phase_c2.txt 2>&1
timeout 300 python tools/phase_times.py C3 4 > gpurun_out/r2a_phase_c3.txt 2>&1
timeout 300 python tools/layer_times.py body25 16 736 1312 > gpurun_out/r2a_layers_body25_16.txt 2>&1
timeout 300 python tools/layer_times.py coco 16 736 984 > gpurun_out/r2a_layers_coco_16.txt 2>&1
timeout 300 python tools/layer_times.py hand 32 736 736 > gpurun_out/r2a_layers_hand_32.txt 2>&1
timeout 300 python tools/layer_times.py coco 1 736 984 > gpurun_out/r2a_layers_coco_1.txt 2>&1
timeout 300 python tools/layer_times.py coco 1 184 248 > gpurun_out/r2a_layers_coco_1_small.txt 2>&1
echo done
