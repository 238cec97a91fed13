"""Two chained 1x1 layers in one launch (csrc/conv_umma.cu variant 6; Mconv6 -> Mconv7 etc., src/model.py:57-62) against
torch on the kernels' operand rounding (bf16 inputs / weights / intermediate, float32 accumulation), and whole networks with
the pairs fused against the same networks launched layer by layer."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import isl_b200  # noqa: E402
from isl_b200 import _lib, synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _pack(w):
    L = _lib.lib()
    cout, cin = w.shape[0], w.shape[1]
    out = torch.empty((1, cout, cin), dtype=torch.bfloat16, device=w.device)
    _lib.check(L.islpose_pack_conv_weights(_lib.ptr(w.contiguous()), cout, cin, 1, None, cin, cin, 0, _lib.ptr(out), _lib.stream_ptr()), "pack")
    return out


@pytest.mark.parametrize("cfg", [(2, 23, 31, 128, 128, 38, "none"), (1, 46, 62, 128, 512, 19, "relu"), (1, 40, 52, 288, 256, 52, "prelu"),
                                 (2, 17, 29, 384, 512, 26, "prelu"), (1, 9, 7, 128, 128, 22, "none"), (1, 33, 40, 64, 64, 64, "relu")])
def test_pair_equals_two_convolutions_on_bf16_operands(cfg):
    n, h, w, cin, mid, cout2, act1 = cfg
    L = _lib.lib()
    dev = torch.device("cuda")
    g = torch.Generator().manual_seed(cin + mid + cout2)
    cs = cin + 64                                   # input slice inside a wider buffer
    xin = (torch.randn((n, h, w, cs), generator=g) * 0.5).to(torch.bfloat16).to(dev)
    w1 = (torch.randn((mid, cin, 1, 1), generator=g) * (1.5 / cin ** 0.5)).to(dev)
    w2 = (torch.randn((cout2, mid, 1, 1), generator=g) * (1.5 / mid ** 0.5)).to(dev)
    b1 = torch.zeros(512, device=dev)
    b1[:mid] = (torch.randn((mid,), generator=g) * 0.2).to(dev)
    b2 = torch.zeros(512, device=dev)
    b2[:cout2] = (torch.randn((cout2,), generator=g) * 0.2).to(dev)
    s1 = torch.zeros(512, device=dev)
    if act1 == "none":
        s1[:mid] = 1.0
    elif act1 == "prelu":
        s1[:mid] = (torch.rand((mid,), generator=g) * 0.5).to(dev)
    s2 = torch.zeros(512, device=dev)
    s2[:cout2] = 1.0                                # the second layer of every pair has no activation (model.py:215-218)
    p1, p2 = _pack(w1), _pack(w2)
    c2s = (cout2 + 7) // 8 * 8
    out_a = torch.full((n, h, w, c2s + 16), 3.0, dtype=torch.bfloat16, device=dev)
    out_b = torch.full((n, h, w, c2s + 8), 3.0, dtype=torch.bfloat16, device=dev)
    out_f = torch.full((n, cout2 + 2, h, w), 5.0, dtype=torch.float32, device=dev)
    d = _lib.ConvDesc()
    d.in_ = xin.data_ptr()
    d.in_c, d.in_cstride, d.in_c_readable, d.w_cin = cin, cs, cs, cin
    d.n, d.h, d.w = n, h, w
    d.weights, d.cout, d.ksize = p1.data_ptr(), mid, 1
    d.bias, d.slope = b1.data_ptr(), s1.data_ptr()
    d.weights2, d.cout2, d.bias2, d.slope2 = p2.data_ptr(), cout2, b2.data_ptr(), s2.data_ptr()
    d.out2_bf16, d.out2_cstride = out_a.data_ptr() + 2 * 8, out_a.shape[3]
    d.out2b_bf16, d.out2b_cstride = out_b.data_ptr(), out_b.shape[3]
    d.out2_f32, d.out2_f32_channels = out_f.data_ptr(), out_f.shape[1]
    plan = C.c_void_p()
    _lib.check(L.islpose_plan_create(C.byref(plan)), "create")
    try:
        _lib.check(L.islpose_plan_add_conv(plan, C.byref(d)), "add_conv(pair)")
        _lib.check(L.islpose_plan_set_graph(plan, 0), "set_graph")
        _lib.check(L.islpose_plan_run(plan, _lib.stream_ptr()), "run")
        torch.cuda.synchronize()
    finally:
        L.islpose_plan_destroy(plan)
    x = xin[..., :cin].to(torch.float64).permute(0, 3, 1, 2)
    y1 = torch.nn.functional.conv2d(x, w1.to(torch.bfloat16).to(torch.float64), b1[:mid].to(torch.float64))
    y1 = torch.where(y1 > 0, y1, y1 * s1[:mid].to(torch.float64).view(1, -1, 1, 1)).to(torch.float32).to(torch.bfloat16).to(torch.float64)
    y2 = torch.nn.functional.conv2d(y1, w2.to(torch.bfloat16).to(torch.float64), b2[:cout2].to(torch.float64))
    ref = y2.to(torch.float32)
    got = out_f[:, :cout2]
    # float32 accumulation in another order; an intermediate value that rounds to the other bf16 neighbour moves the sum a little
    tol = 2e-3 * float(ref.abs().max())
    assert float((got - ref).abs().max()) <= tol, (float((got - ref).abs().max()), tol)
    assert bool((out_f[:, cout2:] == 5.0).all())
    a = out_a[..., 8:8 + c2s].to(torch.float32).permute(0, 3, 1, 2)
    assert torch.equal(a[:, :cout2], got.to(torch.bfloat16).to(torch.float32))     # the bf16 slice is the rounded float32 result
    assert bool((a[:, cout2:] == 0).all())                                          # pad channels are exact zeros
    assert torch.equal(out_b[..., :c2s], out_a[..., 8:8 + c2s])
    assert bool((out_a[..., :8] == 3.0).all()) and bool((out_a[..., 8 + c2s:] == 3.0).all()) and bool((out_b[..., c2s:] == 3.0).all())


@pytest.mark.parametrize("kind,shape", [("coco", (2, 184, 248)), ("hand", (3, 96, 96)), ("body25", (1, 184, 328))])
def test_networks_with_fused_pairs_equal_layer_by_layer(kind, shape):
    n, h, w = shape
    flat = synth.make_flat_weights(kind, seed=3, init="he")
    fused = isl_b200.PoseNet(kind, flat)
    plain = isl_b200.PoseNet(kind, flat, tuning={"pair": False})
    names = fused.instance(n, h, w).op_names
    assert sum("+" in nm and "pool" not in nm for nm in names) >= 6, names
    assert all("+Mconv7" not in nm and "+conv" not in nm for nm in plain.instance(n, h, w).op_names)
    x = (torch.rand((n, 3, h, w), generator=torch.Generator().manual_seed(1)) - 0.5).cuda()
    a = fused(x)
    b = plain(x)
    a = a if isinstance(a, (tuple, list)) else (a,)
    b = b if isinstance(b, (tuple, list)) else (b,)
    for u, v in zip(a, b):
        assert torch.equal(u, v), float((u - v).abs().max())
