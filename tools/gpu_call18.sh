#!/bin/bash
# round 2, GPU call 18 (8 GPUs): the bench under torchrun on eight ranks with the final build; gloo tests are CPU-side
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2q
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 8 --warmup 3 --sub-steps 3 > gpurun_out/${T}_bench_8gpu.json 2> gpurun_out/${T}_bench_8gpu.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${T}_bench_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2q_bench_8gpu.json').read().strip().splitlines()[-1])
print('C2', d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['frac'])
for s in d['sub_results']:
    print(s['config']['workload'][:2], s['value'], s['e2e']['value'], s['roofline']['frac'])
PY
echo done
