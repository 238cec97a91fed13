#!/bin/bash
# round 2, GPU call 13: grouping statistics, smoke with the classifier, bench's classifier leg, ncu of the new first layer
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out /tmp/ncu
T=r2l
timeout 300 python tools/debug_group_stats.py C2 > gpurun_out/${T}_group_stats_c2.txt 2>&1
cut -c1-200 gpurun_out/${T}_group_stats_c2.txt | tail -24
timeout 300 python tools/debug_group_stats.py C3 > gpurun_out/${T}_group_stats_c3.txt 2>&1
cut -c1-200 gpurun_out/${T}_group_stats_c3.txt | tail -28
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.txt 2>&1
tail -2 gpurun_out/${T}_smoke.txt
timeout 600 python bench.py --no-sub --no-cpu-baseline --steps 4 --warmup 3 > gpurun_out/${T}_bench_head.json 2> gpurun_out/${T}_bench_head.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${T}_bench_head.err
python -c "
import json;d=json.loads(open('gpurun_out/${T}_bench_head.json').read().strip().splitlines()[-1]);print(d['value'],d['e2e'],d['classifier'],d['config'].get('single_frame_latency_ms'))"
timeout 300 ncu --set full --clock-control none -k regex:conv_first -c 2 -o /tmp/ncu/${T}_first python tools/layer_times.py coco 16 736 984 > gpurun_out/${T}_ncu_first.log 2>&1
python tools/ncu_summary.py /tmp/ncu/${T}_first.ncu-rep > gpurun_out/${T}_ncu_first.txt 2>&1
grep -E "duration|occupancy|issue active|dram|L1/TEX|stall|lsu|alu" gpurun_out/${T}_ncu_first.txt | head -40
echo done
