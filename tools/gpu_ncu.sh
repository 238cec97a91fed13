#!/bin/bash
# round 2: ncu captures of the shipped kernels (one gpurun call; every profiled command first runs plain). The reports are
# summarised on the box (tools/ncu_summary.py, raw CSV pages) and only the two small conv reports travel back: gpurun
# merges at most 64 MiB.
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none"
# (1) the dominant conv kernel: v5 on the 7x7 128->128 CPM layer (92x164x8) and on a short-K body25 dense-block layer
build/conv_test v5 8 > gpurun_out/r2_plain_conv8.log 2>&1 &&
$NCU --import-source on -k regex:halo_swapped -s 3 -c 1 -o gpurun_out/r2_ncu_conv7x7 build/conv_test v5 8 > gpurun_out/r2_ncu_conv7x7.log 2>&1
build/conv_test v5 26 > gpurun_out/r2_plain_conv26.log 2>&1 &&
$NCU --import-source on -k regex:halo_swapped -s 3 -c 1 -o gpurun_out/r2_ncu_conv3x3 build/conv_test v5 26 > gpurun_out/r2_ncu_conv3x3.log 2>&1
# (2) every other kernel of a C2 call, one launch each
python tools/ncu_kernels.py C2 4 > gpurun_out/r2_plain_kernels.log 2>&1 &&
$NCU --profile-from-start off -k regex:'^(?!.*conv_umma).*$' -c 40 -o /tmp/ncu/r2_ncu_kernels python tools/ncu_kernels.py C2 4 > gpurun_out/r2_ncu_kernels.log 2>&1
python tools/ncu_summary.py /tmp/ncu/r2_ncu_kernels.ncu-rep > gpurun_out/r2_ncu_kernels.txt 2>&1
ncu -i /tmp/ncu/r2_ncu_kernels.ncu-rep --page raw --csv > gpurun_out/r2_ncu_kernels_raw.csv 2>/dev/null
# (3) launch list of a short bench run: every launch with its device time
python bench.py --no-sub --no-cpu-baseline --steps 2 --warmup 3 --batch 8 > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/r2_launches_c2_batch8.csv \
    python bench.py --no-sub --no-cpu-baseline --steps 2 --warmup 3 --batch 8 > gpurun_out/r2_ncu_bench.log 2>&1
du -sh gpurun_out
echo done
