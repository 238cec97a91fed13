#!/bin/bash
# round 2, GPU call 5: GPU tests, single-frame phases with SM shares, bench, the video loop from disk
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
timeout 300 python tools/phase_times.py C2 1 > gpurun_out/r2d_phase_c2_1.txt 2>&1
timeout 300 python tools/phase_times.py C3 1 > gpurun_out/r2d_phase_c3_1.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
echo "bench rc=$?"
tail -c 400 gpurun_out/r2d_bench.err
timeout 600 python tools/video_loop.py --frames 480 > gpurun_out/r2d_video_1gpu.json 2> gpurun_out/r2d_video_1gpu.err
timeout 600 python tools/video_loop.py --frames 480 --raw > gpurun_out/r2d_video_raw_1gpu.json 2>> gpurun_out/r2d_video_1gpu.err
cat gpurun_out/r2d_video_1gpu.json gpurun_out/r2d_video_raw_1gpu.json
tail -c 300 gpurun_out/r2d_video_1gpu.err
echo done
