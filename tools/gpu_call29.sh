#!/bin/bash
# round 2, GPU call 29: ncu of the final pair kernel (persistent), first layer and map accumulation (three CTAs per SM)
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
T=r2zb
python tools/ncu_kernels.py C2 4 > gpurun_out/${T}_plain_kernels.log 2>&1 &&
ncu --set full --clock-control none --profile-from-start off -k regex:'pair_kernel|conv_first|resize_accumulate' -c 40 -o /tmp/ncu/${T}_ncu python tools/ncu_kernels.py C2 4 > gpurun_out/${T}_ncu.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${T}_ncu.ncu-rep > gpurun_out/${T}_ncu.txt 2>&1
grep -c "^==" gpurun_out/${T}_ncu.txt
grep -E "^==|duration|grid  |regs|achieved occ|tensor pipe % of elapsed|dram throughput|issue active % \(smsp\)" gpurun_out/${T}_ncu.txt | head -80 | cut -c1-150
echo done
