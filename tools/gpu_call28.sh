#!/bin/bash
# round 2, GPU call 28: gaussian + NMS at three and at four CTAs per SM (80 registers, 256 bytes spilled)
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
cp isl-signlanguage-translation_b200/libislpose.so /tmp/lib_default.so
for v in default g4 default g4; do
  if [ $v != default ]; then cp build/libislpose_$v.so isl-signlanguage-translation_b200/libislpose.so; else cp /tmp/lib_default.so isl-signlanguage-translation_b200/libislpose.so; fi
  for wl in C2 C3; do
    timeout 300 python tools/phase_times.py $wl 4 > gpurun_out/r2za_${v}_$wl.txt 2>&1
    echo "$v $wl: $(grep 'gaussian' gpurun_out/r2za_${v}_$wl.txt) | $(grep 'two pass' gpurun_out/r2za_${v}_$wl.txt | cut -c1-45)"
  done
done
cp /tmp/lib_default.so isl-signlanguage-translation_b200/libislpose.so
echo done
