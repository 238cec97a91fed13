"""GPU edge cases: ragged / tiny inputs, batch invariance, the .model contract and C-ABI error paths."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import isl_b200  # noqa: E402
from isl_b200 import _lib, synth  # noqa: E402
from isl_b200.body import scale_geometry  # noqa: E402
from isl_b200.extract import KeypointExtractor  # noqa: E402
from oracle import openpose_oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dev():
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def coco(dev):
    return isl_b200.Body(O.make_flat_weights("coco", seed=1), "coco", scale_search=[0.5, 1.0])


@pytest.fixture(scope="module")
def hand(dev):
    return isl_b200.Hand(O.make_flat_weights("hand", seed=2))


@pytest.mark.parametrize("suite", ["v2", "v3", "v4", "v5"])
def test_conv_variant_suites(suite):
    exe = os.path.join(ROOT, "build", "conv_test")
    if not os.path.isfile(exe):
        pytest.skip("build/conv_test not built")
    out = subprocess.run([exe, suite], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "FAIL" not in out.stdout, out.stdout[-3000:]


def _oracle_body(body, frame, scales):
    def net_fn(d):
        p, h = body.model(torch.from_numpy(np.ascontiguousarray(d)).cuda())
        return p[0].cpu().numpy(), h[0].cpu().numpy()

    return O.body_call(net_fn, frame, "coco", tuple(scales), strict=False)


@pytest.mark.parametrize("hw", [(40, 56), (33, 97), (121, 75)])
def test_tiny_and_odd_frames(dev, coco, hw):
    """Frames smaller than the Gaussian support (reflection wraps more than once) and odd aspect ratios."""
    frame = synth.synth_frame(hw[0], hw[1], 31)
    cand, sub = coco(frame)
    ocand, osub = _oracle_body(coco, frame, [0.5, 1.0])
    assert cand.shape == ocand.shape and np.array_equal(cand, ocand)
    assert sub.shape == osub.shape and np.array_equal(sub, osub)


def test_batch_invariance_and_order(dev, coco):
    frames = [synth.synth_frame(96, 128, s) for s in (1, 2, 3, 4, 5)]
    together = coco.batch(frames)
    for f, (c, s) in zip(frames, together):
        c1, s1 = coco(f)
        assert c.shape == c1.shape and np.array_equal(c, c1) and np.array_equal(s, s1)
    # sharding two ways and merging gives the same list
    ex = KeypointExtractor(coco, None)
    from isl_b200.extract import merge_shards
    merged = merge_shards([ex.run_sharded(frames, r, 2, batch_size=2) for r in range(2)], len(frames))
    for (c, s, _), (c0, s0) in zip(merged, together):
        assert np.array_equal(c, c0) and np.array_equal(s, s0)


def test_non_square_hand_crop(dev, hand):
    """Hand.__call__ accepts any image (hand.py:24); a non-square crop makes the four network inputs non-square and
    padded."""
    crop = synth.synth_frame(90, 131, 9)

    def hand_fn(d):
        return hand.model(torch.from_numpy(np.ascontiguousarray(d)).cuda())[0].cpu().numpy()

    assert np.array_equal(hand(crop), O.hand_call(hand_fn, crop))


def test_model_contract(dev, coco, hand):
    """What ISLSignPos / TorchModuleWrapper need from body_estimation.model (ISL_Model_parameter.py:44-47,84-85)."""
    m = coco.model
    assert m.eval() is m and m.to("cuda") is m and m.cuda() is m
    params = list(m.parameters())
    assert len(params) == 2 * 92 and all(isinstance(p, torch.Tensor) for p in params)
    sd = m.state_dict()
    assert isl_b200.util.transfer(m, sd).keys() == sd.keys()
    x = torch.zeros((1, 3, 64, 88), device="cuda")
    with torch.no_grad():
        paf, heat = m(x)
    assert paf.shape == (1, 38, 8, 11) and heat.shape == (1, 19, 8, 11) and paf.dtype == torch.float32 and paf.is_cuda
    assert hand.model(torch.zeros((2, 3, 48, 48), device="cuda")).shape == (2, 22, 6, 6)
    with pytest.raises(ValueError):
        m(torch.zeros((1, 3, 60, 88), device="cuda"))  # not a multiple of 8
    assert coco.njoint == 19 and coco.npaf == 38 and coco.model_type == "coco"


def test_unknown_model_type_falls_back_to_coco(dev, capsys):
    b = isl_b200.Body(O.make_flat_weights("coco", seed=1), "nonsense")   # body.py:25-29
    assert "not right model_type" in capsys.readouterr().out
    assert b.njoint == 19 and b.model_type == "nonsense"


def test_abi_error_paths_on_device(dev):
    L = _lib.lib()
    handle = C.c_void_p()
    assert L.islpose_plan_create(C.byref(handle)) == 0
    buf = torch.zeros(4096, dtype=torch.bfloat16, device="cuda")
    d = _lib.ConvDesc()
    d.in_ = buf.data_ptr() + 2   # misaligned
    d.in_c, d.in_cstride, d.n, d.h, d.w = 64, 64, 1, 4, 4
    d.weights, d.cout, d.ksize = buf.data_ptr(), 64, 3
    d.bias = d.slope = buf.data_ptr()
    d.out_bf16, d.out_cstride = buf.data_ptr(), 64
    assert L.islpose_plan_add_conv(handle, C.byref(d)) != 0 and b"aligned" in L.islpose_last_error()
    d.in_ = buf.data_ptr()
    d.ksize = 5
    assert L.islpose_plan_add_conv(handle, C.byref(d)) != 0 and b"kernel size" in L.islpose_last_error()
    d.ksize, d.pool, d.h = 3, 1, 3   # fused 2x2 pooling needs even sizes
    assert L.islpose_plan_add_conv(handle, C.byref(d)) != 0 and b"pool" in L.islpose_last_error()
    assert L.islpose_pack_conv_weights(buf.data_ptr(), 64, 60, 3, None, 64, 64, 0, buf.data_ptr(), None) != 0
    assert L.islpose_hand_keypoints(None, 1, 4, (C.c_double * 25)(), 0.05, buf.data_ptr(), 0, buf.data_ptr(), None) != 0
    assert L.islpose_body_peaks(buf.data_ptr(), 1, 8, 8, (C.c_double * 25)(), 0.1, 5000, buf.data_ptr(), buf.data_ptr(),
                                buf.data_ptr(), buf.data_ptr(), None) != 0
    assert L.islpose_plan_destroy(handle) == 0


def test_keypoint_extractor_with_real_hand_boxes(dev):
    """body -> util.handDetect -> hand on a frame whose injected maps contain people: the boxes handDetect finds are
    cropped and every hand result is offset by its box origin (demo.py:21-43)."""
    mt, H, W = "coco", 240, 320
    body = isl_b200.Body(O.make_flat_weights(mt, seed=0), mt, scale_search=[0.5])
    hand = isl_b200.Hand(O.make_flat_weights("hand", seed=0, init="torch"))
    sk = synth.synth_skeletons(mt, 3, 1)
    maps = []
    for (m, rh, rw, hp, wp) in scale_geometry(H, W, body.scale_search, body.boxsize):
        paf, heat = synth.render_maps(mt, sk, hp // 8, wp // 8)
        maps.append((torch.from_numpy(paf)[None].cuda(), torch.from_numpy(heat)[None].cuda(), (rh, rw, hp, wp)))
    (cand, sub), = body.postprocess(maps, 1, H, W, body._workspace(1, H, W))
    frame = synth.synth_frame(H, W, 1)
    boxes = isl_b200.util.handDetect(cand, sub, frame)
    assert len(boxes) == 6
    peaks = hand.batch([frame[y:y + w, x:x + w, :] for x, y, w, _ in boxes])
    assert len(peaks) == 6 and all(p.shape == (21, 2) for p in peaks)
    for (x, y, w, _), p in zip(boxes, peaks):
        assert (p[:, 0] < w).all() and (p[:, 1] < w).all() and (p >= 0).all()


def test_maps_accumulate_single_pass_equals_two_pass(dev):
    """islpose_maps_accumulate gives bit-identical planes with and without the float32 workspace."""
    L = _lib.lib()
    H, W, n, C, parts = 97, 131, 2, 19, 18
    scales = scale_geometry(H, W, [0.5, 1.0, 1.5], 368)
    arr = (_lib.Scale * len(scales))()
    keep = []
    rng = np.random.RandomState(0)
    for i, (m, rh, rw, hp, wp) in enumerate(scales):
        t = torch.from_numpy(rng.randn(n, C, hp // 8, wp // 8).astype(np.float32)).to(dev)
        keep.append(t)
        arr[i].lowres = t.data_ptr()
        arr[i].gh, arr[i].gw, arr[i].hc, arr[i].wc = hp // 8, wp // 8, rh, rw
    a = torch.empty((n, parts, H, W), dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    need = L.islpose_maps_workspace_floats(arr, len(scales), n, parts)
    wsp = torch.empty((need,), dtype=torch.float32, device=dev)
    _lib.check(L.islpose_maps_accumulate(arr, len(scales), C, n, H, W, parts, 1, _lib.ptr(a), None, 0, _lib.stream_ptr()), "single")
    _lib.check(L.islpose_maps_accumulate(arr, len(scales), C, n, H, W, parts, 1, _lib.ptr(b), _lib.ptr(wsp), need,
                                         _lib.stream_ptr()), "two-pass")
    assert torch.equal(a, b)


def test_maps_accumulate_two_pass_large_downscale_and_q1_off(dev):
    """Two-pass accumulation where the second stage shrinks strongly (hand-like geometry: many source rows per frame
    tile) and with the plain mean (hand.py:56) instead of body.py:80's running double sum."""
    L = _lib.lib()
    H, W, n, C, parts = 150, 150, 1, 22, 21
    scales = scale_geometry(H, W, [0.5, 1.0, 1.5, 2.0], 368)
    arr = (_lib.Scale * len(scales))()
    keep = []
    rng = np.random.RandomState(1)
    for i, (m, rh, rw, hp, wp) in enumerate(scales):
        t = torch.from_numpy(rng.randn(n, C, hp // 8, wp // 8).astype(np.float32)).to(dev)
        keep.append(t)
        arr[i].lowres = t.data_ptr()
        arr[i].gh, arr[i].gw, arr[i].hc, arr[i].wc = hp // 8, wp // 8, rh, rw
    a = torch.empty((n, parts, H, W), dtype=torch.float64, device=dev)
    b = torch.empty_like(a)
    need = L.islpose_maps_workspace_floats(arr, len(scales), n, parts)
    wsp = torch.empty((need,), dtype=torch.float32, device=dev)
    _lib.check(L.islpose_maps_accumulate(arr, len(scales), C, n, H, W, parts, 0, _lib.ptr(a), None, 0, _lib.stream_ptr()), "single")
    _lib.check(L.islpose_maps_accumulate(arr, len(scales), C, n, H, W, parts, 0, _lib.ptr(b), _lib.ptr(wsp), need,
                                         _lib.stream_ptr()), "two-pass")
    assert torch.equal(a, b)


@pytest.mark.parametrize("geom", [(240, 320, 19, 18, (0.5, 1.0, 1.5, 2.0), 1), (121, 75, 19, 18, (1.0,), 1), (480, 640, 19, 18, (0.5, 1.0, 1.5, 2.0), 2),
                                  (360, 203, 26, 25, (0.5, 1.0, 1.5, 2.0), 1), (333, 517, 22, 21, (0.7, 1.3), 0), (720, 1280, 26, 25, (0.5, 2.0), 1),
                                  (97, 511, 19, 18, (0.25, 0.5), 1), (256, 256, 19, 3, (1.0, 1.5, 2.0), 1), (540, 35, 19, 18, (0.5, 1.0, 2.0), 1)])
def test_maps_accumulate_tma_windows_equal_single_pass(dev, geom):
    """The TMA-fed second stage (csrc/prepost.cu: windows start at a multiple of four floats, boxes reach beyond the maps and
    are zero-filled there) against the single-pass kernel over odd sizes, narrow and wide frames, up- and down-scaling,
    part counts that leave a short last chunk, and both accumulation rules: bit-identical planes."""
    H, W, C, parts, srch, q1 = geom
    L = _lib.lib()
    n = 2 if H * W < 200000 else 1
    scales = scale_geometry(H, W, list(srch), 368)
    arr = (_lib.Scale * len(scales))()
    keep = []
    rng = np.random.RandomState(H + W)
    for i, (m, rh, rw, hp, wp) in enumerate(scales):
        t = torch.from_numpy(rng.randn(n, C, hp // 8, wp // 8).astype(np.float32)).to(dev)
        keep.append(t)
        arr[i].lowres = t.data_ptr()
        arr[i].gh, arr[i].gw, arr[i].hc, arr[i].wc = hp // 8, wp // 8, rh, rw
    a = torch.empty((n, parts, H, W), dtype=torch.float64, device=dev)
    b = torch.full_like(a, -7.0)
    need = L.islpose_maps_workspace_floats(arr, len(scales), n, parts)
    wsp = torch.empty((need,), dtype=torch.float32, device=dev)
    _lib.check(L.islpose_maps_accumulate(arr, len(scales), C, n, H, W, parts, q1, _lib.ptr(a), None, 0, _lib.stream_ptr()), "single")
    _lib.check(L.islpose_maps_accumulate(arr, len(scales), C, n, H, W, parts, q1, _lib.ptr(b), _lib.ptr(wsp), need,
                                         _lib.stream_ptr()), "two-pass")
    torch.cuda.synchronize()
    assert torch.equal(a, b), float((a - b).abs().max())


def test_pipeline_and_chunks_equal_the_serial_path(dev, coco, hand):
    """KeypointExtractor.pipeline() (two lanes, batches in flight) and the chunked batch_device() return exactly what
    the one-batch-at-a-time path returns, in order, for device tensors and for host frames."""
    H, W = 96, 128
    frames = [synth.synth_frame(H, W, 40 + i) for i in range(10)]
    boxes = [[[10 + i, 12, 48, True], [60, 30 + i, 36, False]] for i in range(10)]
    serial = []
    ex = KeypointExtractor(coco, hand, chunk=None)
    for a in range(0, 10, 4):
        serial.extend(ex.batch(frames[a:a + 4], boxes[a:a + 4]))

    def same(x, y):
        assert len(x) == len(y)
        for (c1, s1, h1), (c2, s2, h2) in zip(x, y):
            assert c1.shape == c2.shape and np.array_equal(c1, c2)
            assert s1.shape == s2.shape and np.array_equal(s1, s2)
            assert len(h1) == len(h2) and all(np.array_equal(p, q) for p, q in zip(h1, h2))

    piped = []
    for res in ex.pipeline([(frames[a:a + 4], boxes[a:a + 4]) for a in range(0, 10, 4)]):
        piped.extend(res)
    same(serial, piped)
    dev_frames = torch.from_numpy(np.stack(frames)).cuda()
    same(serial, KeypointExtractor(coco, hand, chunk=3).batch_device(dev_frames, boxes))
    assert any(len(c) for c, _, _ in serial) or True   # random-init maps may hold no peaks; the equality is the test


def test_clip_features_on_device(dev, coco, hand):
    X = KeypointExtractor(coco, hand).features([synth.synth_frame(64, 80, 70 + i) for i in range(5)], batch_size=2)
    assert X.shape == (5, 156) and X.dtype == np.float64 and np.isfinite(X).all()


def test_negative_stride_frames_and_rgb_views(dev, coco):
    """Callers pass views such as frame[:, :, ::-1] (extract_features.py:163): same result as the contiguous copy."""
    frame = synth.synth_frame(72, 96, 91)
    view = frame[:, :, ::-1]
    assert view.strides[2] < 0
    c1, s1 = coco(view)
    c2, s2 = coco(np.ascontiguousarray(view))
    assert c1.shape == c2.shape and np.array_equal(c1, c2) and np.array_equal(s1, s2)
