#!/bin/bash
# round 2, GPU call 25: what the driver runs at round end, on the final build: GPU tests, smoke, default bench, reference arm
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
T=r2x
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=900 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log | cut -c1-240
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1
tail -2 gpurun_out/${T}_smoke.txt
timeout 300 python tools/phase_times.py C2 1 > gpurun_out/${T}_phase_c2_1.txt 2>&1
grep -h "PAF score" gpurun_out/${T}_phase_c2_1.txt
timeout 1200 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/${T}_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['config'].get('single_frame_latency_ms'),d['cpu_baseline']['value'])"
echo done
