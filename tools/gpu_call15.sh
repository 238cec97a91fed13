#!/bin/bash
# round 2, GPU call 15: the records of the final build: full bench line, phase timings, ncu captures
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
T=r2n
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${T}_bench.err
timeout 300 python tools/phase_times.py C2 4 > gpurun_out/${T}_phase_c2.txt 2>&1
timeout 300 python tools/phase_times.py C2 1 > gpurun_out/${T}_phase_c2_1.txt 2>&1
timeout 300 python tools/phase_times.py C3 4 > gpurun_out/${T}_phase_c3.txt 2>&1
timeout 300 python tools/phase_times.py C3 1 > gpurun_out/${T}_phase_c3_1.txt 2>&1
NCU="ncu --set full --clock-control none"
build/conv_test v5 8 > gpurun_out/${T}_plain_conv8.log 2>&1 &&
$NCU --import-source on -k regex:halo_swapped -s 3 -c 1 -o gpurun_out/${T}_ncu_conv7x7 build/conv_test v5 8 > gpurun_out/${T}_ncu_conv7x7.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_ncu_conv7x7.ncu-rep > gpurun_out/${T}_ncu_conv7x7.txt 2>&1
python tools/ncu_kernels.py C2 4 > gpurun_out/${T}_plain_kernels.log 2>&1 &&
$NCU --profile-from-start off -k regex:'^(?!.*halo_swapped).*$' -c 60 -o /tmp/ncu/${T}_ncu_kernels python tools/ncu_kernels.py C2 4 > gpurun_out/${T}_ncu_kernels.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${T}_ncu_kernels.ncu-rep > gpurun_out/${T}_ncu_kernels.txt 2>&1
python bench.py --no-sub --no-cpu-baseline --steps 2 --warmup 3 --batch 8 > gpurun_out/${T}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/${T}_launches_c2_batch8.csv \
    python bench.py --no-sub --no-cpu-baseline --steps 2 --warmup 3 --batch 8 > gpurun_out/${T}_ncu_bench.log 2>&1
du -sh gpurun_out
echo done
