#!/bin/bash
# round 2, GPU call 10: whole suite, phase timings and the bench with the TMA-fed accumulation and the classifier
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2i
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log
timeout 300 python tools/phase_times.py C2 4 > gpurun_out/${T}_phase_c2.txt 2>&1
timeout 300 python tools/phase_times.py C2 1 > gpurun_out/${T}_phase_c2_1.txt 2>&1
timeout 300 python tools/phase_times.py C3 4 > gpurun_out/${T}_phase_c3.txt 2>&1
grep -h "maps accumulate\|gaussian" gpurun_out/${T}_phase_c2.txt gpurun_out/${T}_phase_c2_1.txt gpurun_out/${T}_phase_c3.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1
tail -2 gpurun_out/${T}_smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"
tail -c 300 gpurun_out/${T}_bench.err
echo done
