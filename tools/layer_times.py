"""Per-layer time of one network plan, launch by launch (CUDA events, no profiler):

    python tools/layer_times.py coco|body25|hand  n  h  w

Prints every launch with its kernel variant, microseconds and algorithmic TFLOP/s, then the layers grouped by name."""
import collections
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import isl_b200  # noqa: E402
from isl_b200 import _lib, synth  # noqa: E402


def main():
    kind, n, h, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    torch.cuda.set_device(0)
    net = isl_b200.PoseNet(kind, synth.make_flat_weights(kind, seed=0, init="torch"))
    inst = net.instance(n, h, w)
    inst.input.uniform_(-0.5, 0.5)
    inst.run()
    torch.cuda.synchronize()
    nl = inst.launches
    ms = np.zeros(nl, dtype=np.float32)
    fl = np.zeros(nl, dtype=np.float64)
    var = np.zeros(nl, dtype=np.int32)
    _lib.check(_lib.lib().islpose_plan_profile(inst.handle, _lib.stream_ptr(), 10, ms.ctypes.data_as(C.c_void_p),
                                               fl.ctypes.data_as(C.c_void_p), var.ctypes.data_as(C.c_void_p)), "plan_profile")
    names = inst.op_names
    groups = collections.OrderedDict()
    for name, t, f, v in zip(names, ms, fl, var):
        key = name
        for tag in ("_stage", "_CPM_L"):
            if tag in name:
                key = name.split(tag)[0] + tag + "*"
        if name.startswith("Mconv") and name[-2] == "_" and name[-1] in "012":
            key = name.split("_")[0][:5] + "N_" + "dense_" + name[-1]
        g = groups.setdefault(key, [0, 0.0, 0.0, set()])
        g[0] += 1
        g[1] += t
        g[2] += f
        g[3].add(int(v))
    total_ms, total_f = float(ms.sum()), float(fl.sum())
    print("== %s n=%d %dx%d: %d launches, %.3f ms launch by launch, %.1f TFLOP/s" % (kind, n, h, w, nl, total_ms, total_f / total_ms / 1e9))
    for k, (cnt, t, f, vs) in groups.items():
        print("  %-22s x%-3d %9.1f us %5.1f%%  %7.1f TFLOP/s  v%s" % (k, cnt, t * 1e3, 100 * t / total_ms, f / t / 1e9 if t > 0 else 0,
                                                                ",".join(str(v) for v in sorted(vs))))


if __name__ == "__main__":
    main()
