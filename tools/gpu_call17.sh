#!/bin/bash
# round 2, GPU call 17 (2 GPUs): the bench under torchrun on two ranks with the final build; gloo tests are CPU-side
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
T=r2p
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 8 --warmup 3 --sub-steps 3 > gpurun_out/${T}_bench_2gpu.json 2> gpurun_out/${T}_bench_2gpu.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${T}_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench_2gpu.json').read().strip().splitlines()[-1])
print('C2', d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['frac'])
for s in d['sub_results']:
    print(s['config']['workload'][:2], s['value'], s['e2e']['value'], s['roofline']['frac'])
PY
echo done
