#!/bin/bash
# round 2, GPU call 9: what the TMA unit accepts for unswizzled float32 boxes (probe), the accumulate bring-up builds, the classifier tests
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
P=build/tma_probe
{
# rank box0 box1 c0 c1 pitch rows planes promo struct
timeout 60 $P 2 16 16 0 0 256 64 1 1 0
timeout 60 $P 2 16 16 3 5 256 64 1 1 0
timeout 60 $P 2 60 55 0 0 984 736 1 1 0
timeout 60 $P 2 60 55 4 9 984 736 1 1 0
timeout 60 $P 2 60 55 7 9 984 736 1 1 0
timeout 60 $P 3 60 55 0 0 984 736 3 1 0
timeout 60 $P 3 60 55 8 9 984 736 3 1 0
timeout 60 $P 3 60 55 7 9 984 736 3 1 0
timeout 60 $P 3 60 55 7 9 984 736 3 0 0
timeout 60 $P 3 60 55 7 9 984 736 3 1 1
timeout 60 $P 3 60 55 8 9 984 736 3 1 1
timeout 60 $P 3 108 103 7 9 228 368 3 1 1
timeout 60 $P 3 108 103 8 9 228 368 3 1 1
timeout 60 $P 3 60 55 960 700 984 736 3 1 1
timeout 60 $P 3 64 55 8 9 984 736 3 1 1
timeout 60 $P 3 32 55 7 9 984 736 3 1 1
} > gpurun_out/r2h_tma_probe.txt 2>&1
cat gpurun_out/r2h_tma_probe.txt
for v in tma tma_a4 tma_p0; do
  timeout 120 python tools/debug_acc.py 240 320 2 build/libislpose_$v.so > gpurun_out/r2h_acc_$v.txt 2>&1
  echo "$v: $(tail -1 gpurun_out/r2h_acc_$v.txt | cut -c1-200)"
done
timeout 600 python -m pytest tests/test_translate.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/r2h_pytest_translate.log 2>&1
tail -5 gpurun_out/r2h_pytest_translate.log
echo done
