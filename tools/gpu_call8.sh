#!/bin/bash
# round 2, GPU call 8: locate the fault of the TMA-fed accumulation (sanitizer), then the whole suite on the build without it
set +e
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 python tools/debug_acc.py 240 320 2 > gpurun_out/r2g_acc_plain.txt 2>&1
echo "plain rc=$?"; tail -5 gpurun_out/r2g_acc_plain.txt | cut -c1-300
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/debug_acc.py 121 75 1 > gpurun_out/r2g_acc_sanitizer.txt 2>&1
echo "sanitizer rc=$?"; grep -v "^=========     at\|^=========         Host" gpurun_out/r2g_acc_sanitizer.txt | head -60 | cut -c1-300
cp build/libislpose_notma.so isl-signlanguage-translation_b200/libislpose.so
timeout 120 python tools/debug_acc.py 240 320 2 > gpurun_out/r2g_acc_notma.txt 2>&1
tail -2 gpurun_out/r2g_acc_notma.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout=900 -p no:cacheprovider > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -8 gpurun_out/r2g_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.txt 2>&1
tail -2 gpurun_out/r2g_smoke.txt
timeout 300 python tools/phase_times.py C2 4 > gpurun_out/r2g_phase_c2.txt 2>&1
timeout 300 python tools/phase_times.py C2 1 > gpurun_out/r2g_phase_c2_1.txt 2>&1
timeout 300 python tools/phase_times.py C3 4 > gpurun_out/r2g_phase_c3.txt 2>&1
timeout 300 python tools/layer_times.py coco 1 184 248 > gpurun_out/r2g_layers_coco_1_small.txt 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"
tail -c 300 gpurun_out/r2g_bench.err
echo done
