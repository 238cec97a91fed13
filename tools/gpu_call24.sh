#!/bin/bash
# round 2, GPU call 24: pair kernel at 3 CTAs per SM (63 registers): heuristic / always persistent / one tile per CTA
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cp isl-signlanguage-translation_b200/libislpose.so /tmp/lib_default.so
for v in default pm1 pm2; do
  if [ $v != default ]; then cp build/libislpose_$v.so isl-signlanguage-translation_b200/libislpose.so; fi
  for cfg in "coco 16 736 984" "body25 16 736 1312" "hand 32 736 736"; do
    set -- $cfg
    timeout 300 python tools/layer_times.py $1 $2 $3 $4 > gpurun_out/r2w_${v}_$1.txt 2>&1
    echo "$v $1: $(head -1 gpurun_out/r2w_${v}_$1.txt | cut -c1-100)"; grep -h "v6" gpurun_out/r2w_${v}_$1.txt
  done
done
cp /tmp/lib_default.so isl-signlanguage-translation_b200/libislpose.so
timeout 600 python -m pytest tests/test_gpu_pair1x1.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -3
echo done
