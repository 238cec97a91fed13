#!/bin/bash
# round 2, GPU call 16: list-based matching rounds: whole suite, phase timings, determinism over repeated calls
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2o
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -8 gpurun_out/${T}_pytest.log | cut -c1-240
timeout 300 python tools/phase_times.py C2 1 > gpurun_out/${T}_phase_c2_1.txt 2>&1
timeout 300 python tools/phase_times.py C2 4 > gpurun_out/${T}_phase_c2.txt 2>&1
grep -h "PAF score" gpurun_out/${T}_phase_c2_1.txt gpurun_out/${T}_phase_c2.txt
timeout 300 python tools/debug_group_stats.py C2 > gpurun_out/${T}_group_stats_c2.txt 2>&1
grep "limb  3\|limb 10\|limb 11\|limb 13\|total" gpurun_out/${T}_group_stats_c2.txt
timeout 600 python bench.py --no-sub --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/${T}_bench_head.json 2> gpurun_out/${T}_bench_head.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_head.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e']['value'],d['config'].get('single_frame_latency_ms'))"
echo done
