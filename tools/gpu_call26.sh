#!/bin/bash
# round 2, GPU call 26 (2 GPUs): both arms under torchrun exactly as the driver launches them, final build
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2y_ref_2gpu.json 2> gpurun_out/r2y_ref_2gpu.err
echo "reference arm rc=$?"; grep -c '"impl": "reference"' gpurun_out/r2y_ref_2gpu.json; cut -c1-200 gpurun_out/r2y_ref_2gpu.json | tail -2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2y_bench_2gpu.json 2> gpurun_out/r2y_bench_2gpu.err
echo "bench rc=$?"; tail -c 200 gpurun_out/r2y_bench_2gpu.err
python - <<'PY'
import json
lines=[l for l in open('gpurun_out/r2y_bench_2gpu.json').read().strip().splitlines() if l.startswith('{')]
print(len(lines), 'json line(s) on stdout')
d=json.loads(lines[-1])
print('C2', d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['frac'], d.get('cpu_baseline'))
for s in d['sub_results']:
    print(s['config']['workload'][:2], s['value'], s['e2e']['value'])
PY
echo done
