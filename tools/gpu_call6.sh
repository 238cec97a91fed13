#!/bin/bash
# round 2, GPU call 6 (2 GPUs): bench under torchrun (C4's host gather over NCCL), the video loop on two ranks
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 --sub-steps 3 > gpurun_out/r2e_bench_2gpu.json 2> gpurun_out/r2e_bench_2gpu.err
echo "bench rc=$?"
tail -c 600 gpurun_out/r2e_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/video_loop.py --frames 960 > gpurun_out/r2e_video_2gpu.json 2> gpurun_out/r2e_video_2gpu.err
cat gpurun_out/r2e_video_2gpu.json
tail -c 300 gpurun_out/r2e_video_2gpu.err
echo done
