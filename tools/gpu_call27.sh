#!/bin/bash
# round 2, GPU call 27: map accumulation at two and at three CTAs per SM (80 registers, 64 bytes spilled)
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cp isl-signlanguage-translation_b200/libislpose.so /tmp/lib_default.so
for v in default lb3 default lb3; do
  if [ $v != default ]; then cp build/libislpose_$v.so isl-signlanguage-translation_b200/libislpose.so; else cp /tmp/lib_default.so isl-signlanguage-translation_b200/libislpose.so; fi
  for wl in C2 C3; do
    timeout 300 python tools/phase_times.py $wl 4 > gpurun_out/r2z_${v}_$wl.txt 2>&1
    echo "$v $wl: $(grep 'two pass' gpurun_out/r2z_${v}_$wl.txt)"
  done
done
cp build/libislpose_lb3.so isl-signlanguage-translation_b200/libislpose.so
timeout 600 python -m pytest tests/test_gpu_edges.py -m gpu -q --timeout=300 -p no:cacheprovider -k "accumulate" 2>&1 | tail -2
cp /tmp/lib_default.so isl-signlanguage-translation_b200/libislpose.so
echo done
