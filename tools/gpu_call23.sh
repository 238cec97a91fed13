#!/bin/bash
# round 2, GPU call 23: pair kernel persistent only where one CTA owns an SM: unit tests, layer timings
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2v
timeout 600 python -m pytest tests/test_gpu_pair1x1.py tests/test_gpu_parity.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/${T}_pytest_pair.log 2>&1
echo "pair rc=$?"; tail -5 gpurun_out/${T}_pytest_pair.log | cut -c1-240
timeout 300 python tools/layer_times.py coco 16 736 984 > gpurun_out/${T}_layers_coco_16.txt 2>&1
timeout 300 python tools/layer_times.py body25 16 736 1312 > gpurun_out/${T}_layers_body25_16.txt 2>&1
timeout 300 python tools/layer_times.py hand 32 736 736 > gpurun_out/${T}_layers_hand_32.txt 2>&1
timeout 300 python tools/layer_times.py coco 1 184 248 > gpurun_out/${T}_layers_coco_1_small.txt 2>&1
head -1 gpurun_out/${T}_layers_coco_16.txt gpurun_out/${T}_layers_body25_16.txt gpurun_out/${T}_layers_hand_32.txt gpurun_out/${T}_layers_coco_1_small.txt
grep -h "v6" gpurun_out/${T}_layers_coco_16.txt gpurun_out/${T}_layers_body25_16.txt gpurun_out/${T}_layers_hand_32.txt
echo done
