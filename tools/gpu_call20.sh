#!/bin/bash
# round 2, GPU call 20: the accumulation's geometry sweep
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_edges.py -m gpu -q --timeout=300 -p no:cacheprovider -k "accumulate" > gpurun_out/r2s_pytest_acc.log 2>&1
tail -12 gpurun_out/r2s_pytest_acc.log | cut -c1-240
echo done
