#!/bin/bash
# round 2, GPU call 12: first layer with a staged halo / folded bias / ReLU convert; grouping statistics; whole suite
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2k
timeout 600 python -m pytest tests/test_gpu_first_conv.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/${T}_pytest_first.log 2>&1
tail -15 gpurun_out/${T}_pytest_first.log | cut -c1-240
timeout 300 python tools/layer_times.py coco 16 736 984 > gpurun_out/${T}_layers_coco_16.txt 2>&1
timeout 300 python tools/layer_times.py body25 16 736 1312 > gpurun_out/${T}_layers_body25_16.txt 2>&1
timeout 300 python tools/layer_times.py hand 32 736 736 > gpurun_out/${T}_layers_hand_32.txt 2>&1
head -3 gpurun_out/${T}_layers_coco_16.txt gpurun_out/${T}_layers_body25_16.txt gpurun_out/${T}_layers_hand_32.txt
timeout 300 python tools/debug_group_stats.py C2 > gpurun_out/${T}_group_stats_c2.txt 2>&1
cat gpurun_out/${T}_group_stats_c2.txt | cut -c1-200 | tail -24
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log | cut -c1-240
timeout 300 python tools/phase_times.py C2 1 > gpurun_out/${T}_phase_c2_1.txt 2>&1
grep -h "PAF score\|body nets" gpurun_out/${T}_phase_c2_1.txt
echo done
