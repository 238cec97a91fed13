"""conv1_1 (csrc/conv_first.cu; src/model.py 'conv1_1' of the three networks) through the C ABI against torch.conv2d on the
kernel's own operand rounding: input and weights rounded to bf16, float32 accumulation, bias, ReLU / PReLU, bf16 result.
Tolerance: one bf16 step of the result (the accumulation order differs; the bias enters as hi + lo bf16 parts)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import isl_b200  # noqa: E402,F401
from isl_b200 import _lib  # noqa: E402

pytestmark = pytest.mark.gpu


def _run_first(x, w, b, slope, relu, cstride=64):
    L = _lib.lib()
    dev = x.device
    n, _, h, wd = x.shape
    wt = torch.empty((1, 64, 32), dtype=torch.bfloat16, device=dev)
    _lib.check(L.islpose_pack_conv_weights(_lib.ptr(w.contiguous()), 64, 3, 3, None, 3, 32, 1, _lib.ptr(wt), _lib.stream_ptr()), "pack")
    bias = torch.zeros(512, dtype=torch.float32, device=dev)
    bias[:64] = b
    sl = torch.zeros(512, dtype=torch.float32, device=dev)
    sl[:64] = slope
    out = torch.full((n, h, wd, cstride), 7.0, dtype=torch.bfloat16, device=dev)
    plan = C.c_void_p()
    _lib.check(L.islpose_plan_create(C.byref(plan)), "create")
    try:
        _lib.check(L.islpose_plan_add_first_conv(plan, _lib.ptr(x), _lib.ptr(wt), _lib.ptr(bias), _lib.ptr(sl), _lib.ptr(out), cstride,
                                                 n, h, wd, 1 if relu else 0), "add_first_conv")
        _lib.check(L.islpose_plan_set_graph(plan, 0), "set_graph")
        _lib.check(L.islpose_plan_run(plan, _lib.stream_ptr()), "run")
        torch.cuda.synchronize()
    finally:
        L.islpose_plan_destroy(plan)
    return out


def _reference(x, w, b, slope):
    xr = x.to(torch.bfloat16).to(torch.float64)
    wr = w.to(torch.bfloat16).to(torch.float64)
    y = torch.nn.functional.conv2d(xr, wr, b.to(torch.float64), padding=1)
    y = torch.where(y > 0, y, y * slope.to(torch.float64).view(1, -1, 1, 1))
    return y.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("shape", [(1, 8, 40), (2, 37, 83), (1, 184, 248), (3, 96, 96), (1, 5, 3)])
@pytest.mark.parametrize("act", ["relu", "prelu"])
def test_first_layer_equals_conv2d_on_bf16_operands(shape, act):
    n, h, w = shape
    g = torch.Generator().manual_seed(h * 1000 + w)
    x = ((torch.rand((n, 3, h, w), generator=g) - 0.5)).cuda()              # the network input range (body.py:55)
    wt = (torch.randn((64, 3, 3, 3), generator=g) * 0.2).cuda()
    b = (torch.randn((64,), generator=g) * 0.3).cuda()
    slope = (torch.zeros(64) if act == "relu" else torch.rand((64,), generator=g) * 0.5).cuda()
    out = _run_first(x, wt, b, slope, relu=(act == "relu"), cstride=72)
    ref = _reference(x, wt, b, slope)
    got = out[..., :64].to(torch.float64)
    step = torch.maximum(ref.abs(), torch.tensor(2.0 ** -6, dtype=torch.float64, device=ref.device)) * 2.0 ** -8   # one bf16 step
    assert bool(((got - ref).abs() <= step + 1e-6).all()), float(((got - ref).abs() / step).max())
    assert bool((out[..., 64:] == 7.0).all())                                # channels beyond the 64 are not touched
    if act == "relu":
        assert bool((got >= 0).all())


def test_first_layer_general_epilogue_equals_the_relu_path():
    """relu = 0 with all-zero slopes takes the general epilogue: same values as the ReLU fast path (-0.0 aside)."""
    g = torch.Generator().manual_seed(5)
    x = (torch.rand((2, 3, 50, 70), generator=g) - 0.5).cuda()
    wt = (torch.randn((64, 3, 3, 3), generator=g) * 0.2).cuda()
    b = (torch.randn((64,), generator=g) * 0.3).cuda()
    z = torch.zeros(64).cuda()
    a = _run_first(x, wt, b, z, relu=True).float()
    c = _run_first(x, wt, b, z, relu=False).float()
    assert torch.equal(a, c)     # float comparison: -0.0 == +0.0
