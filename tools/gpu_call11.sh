#!/bin/bash
# round 2, GPU call 11: grouping kernels with their tables in shared memory: tests, phase timings, ncu of the non-conv kernels
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
T=r2j
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
timeout 300 python tools/phase_times.py C2 1 > gpurun_out/${T}_phase_c2_1.txt 2>&1
timeout 300 python tools/phase_times.py C2 4 > gpurun_out/${T}_phase_c2.txt 2>&1
timeout 300 python tools/phase_times.py C3 1 > gpurun_out/${T}_phase_c3_1.txt 2>&1
grep -h "PAF score\|maps accumulate (two" gpurun_out/${T}_phase_c2_1.txt gpurun_out/${T}_phase_c2.txt gpurun_out/${T}_phase_c3_1.txt
python tools/ncu_kernels.py C2 4 > gpurun_out/${T}_plain_kernels.log 2>&1 &&
ncu --set full --clock-control none --profile-from-start off -k regex:'^(?!.*conv_umma).*$' -c 40 -o /tmp/ncu/${T}_ncu_kernels python tools/ncu_kernels.py C2 4 > gpurun_out/${T}_ncu_kernels.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${T}_ncu_kernels.ncu-rep > gpurun_out/${T}_ncu_kernels.txt 2>&1
grep -c "^==" gpurun_out/${T}_ncu_kernels.txt
echo done
