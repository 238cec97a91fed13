#!/bin/bash
# round 2: ncu captures of the shipped kernels (one gpurun call; every profiled command first runs plain)
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# (1) the dominant conv kernel: v5 on the 7x7 128->128 CPM layer (92x164x8) and on a short-K body25 dense-block layer
build/conv_test v5 9 > gpurun_out/r2_plain_conv9.log 2>&1 &&
$NCU -k regex:halo_swapped -s 3 -c 1 -o gpurun_out/r2_ncu_conv7x7 build/conv_test v5 9 > gpurun_out/r2_ncu_conv7x7.log 2>&1
build/conv_test v5 26 > gpurun_out/r2_plain_conv26.log 2>&1 &&
$NCU -k regex:halo_swapped -s 3 -c 1 -o gpurun_out/r2_ncu_conv3x3 build/conv_test v5 26 > gpurun_out/r2_ncu_conv3x3.log 2>&1
# (2) every other kernel of a C2 call, one launch each
python tools/ncu_kernels.py C2 4 > gpurun_out/r2_plain_kernels.log 2>&1 &&
$NCU --profile-from-start off -k regex:'^(?!.*conv_umma).*$' -c 60 -o gpurun_out/r2_ncu_kernels python tools/ncu_kernels.py C2 4 > gpurun_out/r2_ncu_kernels.log 2>&1
# (3) launch list of a short bench run: every launch with its device time
python bench.py --no-sub --no-cpu-baseline --steps 2 --warmup 3 --batch 8 > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches_c2_batch8.csv \
    python bench.py --no-sub --no-cpu-baseline --steps 2 --warmup 3 --batch 8 > gpurun_out/r2_ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo done
