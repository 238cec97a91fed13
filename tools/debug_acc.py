"""Bring-up check of the two-pass map accumulation (TMA-fed second stage) against the single-pass kernel.
usage: python tools/debug_acc.py [H W n [library]]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from isl_b200 import _lib  # noqa: E402
from isl_b200.body import scale_geometry  # noqa: E402


def main():
    H, W, n = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (240, 320, 2)
    C, parts = 19, 18
    dev = torch.device("cuda:0")
    if len(sys.argv) >= 5:
        _lib.LIB_PATH = os.path.abspath(sys.argv[4])   # a bring-up build of the library
    L = _lib.lib()
    scales = scale_geometry(H, W, [0.5, 1.0, 1.5, 2.0], 368)
    arr = (_lib.Scale * len(scales))()
    keep = []
    rng = np.random.RandomState(0)
    for i, (m, rh, rw, hp, wp) in enumerate(scales):
        t = torch.from_numpy(rng.randn(n, C, hp // 8, wp // 8).astype(np.float32)).to(dev)
        keep.append(t)
        arr[i].lowres = t.data_ptr()
        arr[i].gh, arr[i].gw, arr[i].hc, arr[i].wc = hp // 8, wp // 8, rh, rw
        print("scale", i, "grid", hp // 8, wp // 8, "hc wc", rh, rw, flush=True)
    a = torch.empty((n, parts, H, W), dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    need = L.islpose_maps_workspace_floats(arr, len(scales), n, parts)
    wsp = torch.empty((need,), dtype=torch.float32, device=dev)
    _lib.check(L.islpose_maps_accumulate(arr, len(scales), C, n, H, W, parts, 1, _lib.ptr(a), None, 0, _lib.stream_ptr()), "single")
    torch.cuda.synchronize()
    print("single pass ok", flush=True)
    _lib.check(L.islpose_maps_accumulate(arr, len(scales), C, n, H, W, parts, 1, _lib.ptr(b), _lib.ptr(wsp), need,
                                         _lib.stream_ptr()), "two-pass")
    torch.cuda.synchronize()
    print("two pass ok; equal:", bool(torch.equal(a, b)), "max diff", float((a - b).abs().max()), flush=True)


if __name__ == "__main__":
    main()
