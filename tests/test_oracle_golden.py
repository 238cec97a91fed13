"""The oracle replays every golden vector the reference produced (tests/golden/make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

import isl_b200  # noqa: F401
from isl_b200 import synth
from oracle import openpose_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


@pytest.mark.parametrize("kind", ["coco", "body25", "hand"])
def test_net_forward_matches_reference_module(kind):
    g = _load("net_%s.npz" % kind)
    x = torch.from_numpy(np.random.RandomState(int(g["input_seed"])).uniform(
        -0.5, 0.5, (1, 3, int(g["h"]), int(g["w"]))).astype(np.float32))
    out = O.net_forward(kind, O.make_flat_weights(kind, seed=int(g["weight_seed"])), x)
    outs = [out] if kind == "hand" else list(out)
    for i, o in enumerate(outs):
        ref = g["out%d" % i]
        assert o.shape == ref.shape
        # same torch kernels, same weights: only thread-partitioning differences are tolerated
        np.testing.assert_allclose(o.numpy(), ref, rtol=0, atol=2e-6)


BODY = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "body_*.npz")))


@pytest.mark.parametrize("fixture", BODY)
@pytest.mark.parametrize("backend", ["lib", "restated"])
def test_body_injected_maps_bit_exact(fixture, backend):
    g = _load(fixture)
    mt = str(g["model_type"])
    if backend == "restated" and int(g["h"]) * int(g["w"]) * len(g["scales"]) > 600000:
        pytest.skip("restated backend is covered on the smaller fixtures; this one is slow in numpy")
    sk = synth.synth_skeletons(mt, int(g["people"]), int(g["seed"]))
    drop = set(map(tuple, g["drop"].tolist()))
    img = synth.synth_frame(int(g["h"]), int(g["w"]), int(g["seed"]))

    def fn(data):
        return synth.render_maps(mt, sk, data.shape[2] // 8, data.shape[3] // 8, drop=drop)

    cand, sub = O.body_call(fn, img, mt, tuple(g["scales"].tolist()), backend=backend)
    assert cand.shape == g["candidate"].shape and sub.shape == g["subset"].shape
    assert np.array_equal(cand, g["candidate"])  # float64 scores included: bit exact
    assert np.array_equal(sub, g["subset"])
    boxes = O.hand_detect(cand, sub, img.shape) if len(sub) else []
    got = np.array([[b[0], b[1], b[2], int(b[3])] for b in boxes], dtype=np.int64).reshape(-1, 4)
    assert np.array_equal(got, g["boxes"])


HAND = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLD, "hand_*.npz")))


@pytest.mark.parametrize("fixture", HAND)
@pytest.mark.parametrize("backend", ["lib", "restated"])
def test_hand_injected_maps_bit_exact(fixture, backend):
    g = _load(fixture)
    w, seed = int(g["w"]), int(g["seed"])
    pts = np.random.RandomState(seed).uniform(0.1, 0.9, (21, 2))
    for j in g["missing"].tolist():
        pts[j] = -1
    crop = synth.synth_frame(w, w, seed)
    peaks = O.hand_call(lambda d: synth.render_hand_maps(pts, d.shape[2] // 8, d.shape[3] // 8), crop, backend=backend)
    assert np.array_equal(peaks, g["peaks"])


@pytest.mark.parametrize("fixture", ["bodynet_coco_realnet.npz"])
def test_body_real_network_end_to_end(fixture):
    """Reference Body.__call__ with the real (seeded) network: the oracle in `lib` mode makes the same cv2 /
    scipy / torch calls, so candidate and subset must come out identical."""
    g = _load(fixture)
    mt = str(g["model_type"])
    flat = O.make_flat_weights(mt, seed=int(g["weight_seed"]), gain=float(g["gain"]), head_gain=float(g["head_gain"]))
    img = synth.synth_frame(int(g["h"]), int(g["w"]), int(g["frame_seed"]))
    cand, sub = O.body_call(O.make_net_fn(mt, flat), img, mt, tuple(g["scales"].tolist()), backend="lib")
    assert cand.shape == g["candidate"].shape and sub.shape == g["subset"].shape
    assert np.array_equal(cand[:, [0, 1, 3]], g["candidate"][:, [0, 1, 3]])
    np.testing.assert_allclose(cand[:, 2], g["candidate"][:, 2], rtol=0, atol=1e-6)
    assert np.array_equal(sub[:, :-2], g["subset"][:, :-2])
    np.testing.assert_allclose(sub[:, -2:], g["subset"][:, -2:], rtol=0, atol=1e-5)
