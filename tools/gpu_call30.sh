#!/bin/bash
# round 2, GPU call 30: last check of HEAD: GPU tests and smoke
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout=900 -p no:cacheprovider > gpurun_out/r2zc_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2zc_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo done
