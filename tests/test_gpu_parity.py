"""GPU parity: every stage of the CUDA path (called through the C ABI) against the oracle on the same inputs.

Tolerances
  * integer / index work (resized network input, peaks, candidates, subsets, hand boxes and hand peaks) and the
    float32/float64 map post-processing given identical network outputs: bit exact;
  * network outputs (bf16 operands, fp32 accumulate vs the fp32 reference): max |diff| <= NET_TOL * max |ref|.
"""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import isl_b200  # noqa: E402
from isl_b200 import _lib, synth  # noqa: E402
from isl_b200.body import scale_geometry  # noqa: E402
from oracle import openpose_oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
NET_TOL = 5e-2   # bf16 activations through up to 50 layers, relative to the largest reference value
NET_MEAN_TOL = 1e-2


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "these tests need the B200"
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


def test_conv_bringup_harness():
    exe = os.path.join(ROOT, "build", "conv_test")
    if not os.path.isfile(exe):
        pytest.skip("build/conv_test not built")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ALL PASS" in out.stdout, out.stdout[-3000:]


@pytest.mark.parametrize("hw", [(120, 160), (310, 458), (109, 109), (57, 91)])
@pytest.mark.parametrize("s", [0.5, 1.0, 2.0])
def test_resize_pad_normalize_bit_exact(dev, hw, s):
    h, w = hw
    frames = np.stack([synth.synth_frame(h, w, 3), synth.synth_frame(h, w, 4)])
    (m, rh, rw, hp, wp), = scale_geometry(h, w, [s], 368)
    fd = torch.from_numpy(frames).to(dev)
    out = torch.empty((2, 3, hp, wp), dtype=torch.float32, device=dev)
    u8 = torch.empty((2, hp, wp, 3), dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().islpose_resize_pad_normalize(_lib.ptr(fd), 2, h, w, m, rh, rw, hp, wp, _lib.ptr(out), _lib.ptr(u8),
                                                      _lib.stream_ptr()), "resize")
    for i in range(2):
        data, pshape, pad = O.preprocess(frames[i], m)
        assert data.shape == (1, 3, hp, wp)
        assert np.array_equal(out[i].cpu().numpy(), data[0])
        small = O.resize_cubic(frames[i], fx=m, fy=m)
        assert np.array_equal(u8[i, :rh, :rw].cpu().numpy(), small)


@pytest.mark.parametrize("kind,h,w,n", [("coco", 64, 88, 2), ("body25", 48, 72, 1), ("hand", 96, 96, 2), ("coco", 184, 248, 1)])
def test_network_forward_within_bf16_tolerance(dev, kind, h, w, n):
    flat = O.make_flat_weights(kind, seed=3)
    net = isl_b200.PoseNet(kind, flat)
    x = torch.from_numpy(np.random.RandomState(5).uniform(-0.5, 0.5, (n, 3, h, w)).astype(np.float32))
    ref = O.net_forward(kind, flat, x)
    got = net(x.to(dev))
    refs = [ref] if kind == "hand" else list(ref)
    gots = [got] if kind == "hand" else list(got)
    for r, g in zip(refs, gots):
        r = r.numpy()
        g = g.cpu().numpy()
        assert g.shape == r.shape
        scale = np.abs(r).max()
        assert np.abs(g - r).max() <= NET_TOL * scale, (kind, np.abs(g - r).max() / scale)
        assert np.abs(g - r).mean() <= NET_MEAN_TOL * scale
    if kind == "coco":
        assert gots[1].min().item() >= 0.0  # the final heat map keeps its ReLU (model.py:215-218)


def _inject(body, mt, sk, drop, n, H, W, dev):
    """Per-scale injected network outputs in the layout Body.postprocess expects."""
    maps = []
    for (m, rh, rw, hp, wp) in scale_geometry(H, W, body.scale_search, body.boxsize):
        paf, heat = synth.render_maps(mt, sk, hp // 8, wp // 8, drop=drop)
        maps.append((torch.from_numpy(paf)[None].repeat(n, 1, 1, 1).contiguous().to(dev),
                     torch.from_numpy(heat)[None].repeat(n, 1, 1, 1).contiguous().to(dev), (rh, rw, hp, wp)))
    return maps


BODY = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "body_*.npz")))


@pytest.fixture(scope="module")
def bodies(dev):
    cache = {}

    def get(mt, scales):
        key = (mt, tuple(scales))
        if key not in cache:
            cache[key] = isl_b200.Body(O.make_flat_weights(mt, seed=0), mt, scale_search=list(scales))
        return cache[key]

    return get


@pytest.mark.parametrize("fixture", BODY)
def test_body_postprocess_matches_reference_golden(dev, bodies, fixture):
    """Injected maps -> CUDA accumulate / gaussian+NMS / lazy PAF scoring / grouping == what the reference produced."""
    g = np.load(os.path.join(GOLD, fixture))
    mt = str(g["model_type"])
    H, W = int(g["h"]), int(g["w"])
    body = bodies(mt, g["scales"].tolist())
    sk = synth.synth_skeletons(mt, int(g["people"]), int(g["seed"]))
    drop = set(map(tuple, g["drop"].tolist()))
    n = 2
    ws = body._workspace(n, H, W)
    res = body.postprocess(_inject(body, mt, sk, drop, n, H, W, dev), n, H, W, ws)
    for cand, sub in res:
        assert cand.shape == g["candidate"].shape and sub.shape == g["subset"].shape
        assert np.array_equal(cand, g["candidate"])
        assert np.array_equal(sub, g["subset"])
        boxes = isl_b200.util.handDetect(cand, sub, np.zeros((H, W, 3), np.uint8)) if len(sub) else []
        got = np.array([[b[0], b[1], b[2], int(b[3])] for b in boxes], dtype=np.int64).reshape(-1, 4)
        assert np.array_equal(got, g["boxes"])


def test_heat_accumulate_and_peaks_bit_exact(dev, bodies):
    mt, H, W = "body25", 187, 251
    body = bodies(mt, [0.5, 1.0, 1.5, 2.0])
    sk = synth.synth_skeletons(mt, 3, 9)
    maps = _inject(body, mt, sk, set(), 1, H, W, dev)
    ws = body._workspace(1, H, W)
    body.postprocess(maps, 1, H, W, ws)

    def fn(data):
        return synth.render_maps(mt, sk, data.shape[2] // 8, data.shape[3] // 8)

    heat_avg, paf_avg = O.body_maps(fn, synth.synth_frame(H, W, 0), mt, (0.5, 1.0, 1.5, 2.0))
    got = ws["heat"][0].cpu().numpy()  # [parts, H, W]
    assert np.array_equal(got, np.transpose(heat_avg[:, :, :25], (2, 0, 1)))
    peaks = O.body_peaks(heat_avg, 26)
    counts = ws["counts"].cpu().numpy()
    keys = ws["keys"].cpu().numpy().view(np.uint32)
    scores = ws["scores"].cpu().numpy()
    for p in range(25):
        assert counts[p] == len(peaks[p])
        for i, (x, y, sc, _) in enumerate(peaks[p]):
            assert keys[p, i] == y * W + x and scores[p, i] == sc


HAND = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "hand_*.npz")))


@pytest.fixture(scope="module")
def hand(dev):
    return isl_b200.Hand(O.make_flat_weights("hand", seed=0))


@pytest.mark.parametrize("fixture", HAND)
def test_hand_postprocess_matches_reference_golden(dev, hand, fixture):
    g = np.load(os.path.join(GOLD, fixture))
    w, seed = int(g["w"]), int(g["seed"])
    pts = np.random.RandomState(seed).uniform(0.1, 0.9, (21, 2))
    for j in g["missing"].tolist():
        pts[j] = -1
    maps = []
    for (m, rh, rw, hp, wp) in scale_geometry(w, w, hand.scale_search, hand.boxsize):
        t = torch.from_numpy(synth.render_hand_maps(pts, hp // 8, wp // 8))[None].contiguous().to(dev)
        maps.append((t, 0, (rh, rw, hp, wp)))
    peaks = hand.postprocess(maps, w, w).cpu().numpy().astype(np.int64)
    assert np.array_equal(peaks, g["peaks"])


def test_body_end_to_end_against_oracle_on_own_maps(dev):
    """Real (seeded) coco network, two scales: the CUDA result must equal the oracle's post-processing of the very
    network outputs the CUDA path produced (pre-processing and all map / peak / grouping stages exact)."""
    flat = O.make_flat_weights("coco", seed=1)
    body = isl_b200.Body(flat, "coco", scale_search=[0.5, 1.0])
    frame = synth.synth_frame(120, 160, 21)
    cand, sub = body(frame[:, :, ::-1][:, :, ::-1])  # a negative-stride view, like extract_features.py:163 passes

    def net_fn(d):
        p, h = body.model(torch.from_numpy(np.ascontiguousarray(d)).to(dev))
        return p[0].cpu().numpy(), h[0].cpu().numpy()

    ocand, osub = O.body_call(net_fn, frame, "coco", (0.5, 1.0), strict=False)
    assert cand.shape == ocand.shape and np.array_equal(cand, ocand)
    assert sub.shape == osub.shape and np.array_equal(sub, osub)
    # batching is transparent: the same frame twice in one batch gives the same answer twice
    r = body.batch([frame, frame])
    for c, s in r:
        assert np.array_equal(c, cand) and np.array_equal(s, sub)


def test_hand_end_to_end_against_oracle_on_own_maps(dev):
    flat = O.make_flat_weights("hand", seed=2)
    hand = isl_b200.Hand(flat)
    crops = [synth.synth_frame(64, 64, 5), synth.synth_frame(109, 109, 6), synth.synth_frame(64, 64, 7)]

    def hand_fn(d):
        return hand.model(torch.from_numpy(np.ascontiguousarray(d)).to(dev))[0].cpu().numpy()

    got = hand.batch(crops)
    for c, p in zip(crops, got):
        assert p.shape == (21, 2) and p.dtype == np.int64
        assert np.array_equal(p, O.hand_call(hand_fn, c))


def test_empty_frame_shapes(dev, bodies):
    """body.py:183: with no peaks candidate is np.array([]) of shape (0,), subset is (0, njoint+1)."""
    body = bodies("coco", [0.5])
    H, W = 200, 264
    maps = _inject(body, "coco", synth.synth_skeletons("coco", 0, 1), set(), 1, H, W, dev)
    (cand, sub), = body.postprocess(maps, 1, H, W, body._workspace(1, H, W))
    assert cand.shape == (0,) and sub.shape == (0, 20)
