#!/bin/bash
# round 2, GPU call 21: the reference arm exactly as the driver launches it (N = 1 and under torchrun with N = 2)
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
time timeout 900 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/r2t_ref_1.json 2> gpurun_out/r2t_ref_1.err
echo "rc=$?"; cat gpurun_out/r2t_ref_1.json | cut -c1-900; tail -3 gpurun_out/r2t_ref_1.err
ls oracle/_ref | head
echo done
