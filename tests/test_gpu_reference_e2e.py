"""GPU end to end against what the UNMODIFIED reference produced with real networks (tests/golden/bodynet_*.npz).

Every other end-to-end test feeds the oracle the product's own network outputs, which proves the post-processing exact
but says nothing about how far the bf16 networks (and the documented <= 1 grey level difference between OpenCV's generic
resize, implemented here, and the IPP path the pip wheel dispatches to) move the final result. Here the reference ran
entirely on its own - fp32 torch-CPU networks, cv2 with IPP, scipy - and the CUDA path runs entirely on its own; the
drift between the two candidate / subset tables is measured and bounded:

  matched       a reference peak and a CUDA peak of the same part at the same pixel (exact match) or within 1 px
  missing/extra reference peaks without a CUDA partner / CUDA peaks without a reference partner
  score drift   |reference score - CUDA score| over the matched peaks (scores are unsmoothed heat values)
  subset        rows compared after mapping candidate ids through the peak matching

north_star: "heatmaps/PAFs within a stated fp tolerance, and peak coordinates, candidate indices and subset assignments
bit-exact whenever the reference's float decisions are not within tolerance of a threshold". What that means for a
random-init network was measured on the CPU with the oracle (bf16 operands emulated, oracle/openpose_oracle.py
net_forward(emulate_bf16=True)): a peak is the decision v >= its 4 neighbours on a sigma-3-smoothed map, and at a local
maximum of such a map the margin v - neighbour is a curvature term far below ANY useful map tolerance - on all four
fixtures not one reference peak has a margin above half the measured sup-norm error of the smoothed map. Every peak of
these fixtures is therefore "within tolerance of a threshold"; what can be asserted is
  (1) the maps: |heat_avg(CUDA) - heat_avg(reference)| <= MAP_TOL * max|heat_avg| on a lattice of the frame;
  (2) the statistics of the peak tables: most reference peaks are found at the identical pixel or within one pixel, with
      nearly equal scores. Bounds = measured figures with margin. Measured on B200 (and, in brackets, predicted by the
      CPU emulation): bodynet_coco_realnet 65 % identical / 89 % within 1 px (66 / 90), gained network 58 % / 88 %
      (58 / 87), the He-initialised body25 network at 1280x720 (7953 peaks of smooth noise) 28 % / 69 %;
      heat_avg within 1.8 % of its maximum on all of them.
bodynet_c2_coco_s4 (BASELINE's C2 frame with nn.Conv2d-default weights) is the degenerate case: its heat maps are flat
to 1e-3 of their value, bf16 rounding noise of that size doubles the number of local maxima (642 -> 1271 in the
emulation) although the maps agree to 1.1e-3 of their maximum; for it only (1) and the score drift are asserted.
With trained weights peaks are isolated blobs like the injected maps of tests/test_gpu_parity.py, where the whole
post-processing chain is bit-exact.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import isl_b200  # noqa: E402
from isl_b200 import synth  # noqa: E402
from oracle import openpose_oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# fixture -> bounds: minimum fraction of reference peaks matched exactly / within 1 px, maximum fraction of unmatched
# peaks on either side, maximum score drift relative to the largest reference score
MAP_TOL = 3e-2   # of max|reference heat_avg|; bf16 operands through up to 50 layers (measured: see parity_drift.json)
BOUNDS = {
    "bodynet_coco_realnet.npz": dict(exact=0.50, near=0.80, unmatched=0.20, score=0.05),
    "bodynet_coco_realnet_gained.npz": dict(exact=0.45, near=0.78, unmatched=0.22, score=0.06),
    "bodynet_c2_coco_s4.npz": dict(exact=None, near=None, unmatched=None, score=0.01),   # flat maps: see above
    "bodynet_c3_body25_s4.npz": dict(exact=0.22, near=0.60, unmatched=0.36, score=0.05),   # 7953 peaks of smooth noise
}


def split_parts(cand, W):
    """candidate rows -> list of per-part arrays. Parts are concatenated in order and every part's peaks are in
    row-major order (body.py:99-107), so a new part starts where y*W+x does not increase."""
    if cand.ndim != 2 or len(cand) == 0:
        return []
    key = cand[:, 1] * W + cand[:, 0]
    cuts = np.flatnonzero(np.diff(key) <= 0) + 1
    return np.split(cand, cuts)


def match_peaks(ref, got, H, W, njoint):
    """Matches peaks part by part (split_parts recovers the parts from the row-major order inside each part; if the two
    tables split into different numbers of parts, all peaks are matched against all peaks)."""
    rp, gp = split_parts(ref, W), split_parts(got, W)
    stats = dict(ref=len(ref), got=len(got), exact=0, near=0, missing=0, extra=0, score_max=0.0, score_mean=0.0)
    id_map = {}
    diffs = []
    # align part lists: with random-init nets every part has peaks in both tables; if the counts of parts differ, fall
    # back to matching all peaks of a table against all peaks of the other (part-agnostic, slightly optimistic)
    pairs = list(zip(rp, gp)) if len(rp) == len(gp) else [(ref, got)]
    for r, g in pairs:
        gmap = {(int(x), int(y)): i for i, (x, y) in enumerate(g[:, :2])}
        used = set()
        for row in r:
            x, y = int(row[0]), int(row[1])
            hit = gmap.get((x, y))
            kind = "exact"
            if hit is None or hit in used:
                hit, kind = None, "near"
                for dy in (-1, 0, 1):
                    for dx in (-1, 0, 1):
                        j = gmap.get((x + dx, y + dy))
                        if j is not None and j not in used:
                            hit = j
                            break
                    if hit is not None:
                        break
            if hit is None:
                stats["missing"] += 1
                continue
            used.add(hit)
            stats[kind] += 1
            id_map[int(row[3])] = int(g[hit, 3])
            diffs.append(abs(row[2] - g[hit, 2]))
        stats["extra"] += len(g) - len(used)
    if diffs:
        stats["score_max"] = float(np.max(diffs))
        stats["score_mean"] = float(np.mean(diffs))
    return stats, id_map


def compare_subsets(ref_sub, got_sub, id_map, njoint):
    """Rows of the reference subset, with candidate ids mapped into the CUDA table, looked up among the CUDA rows."""
    got_rows = {tuple(int(v) for v in row[:njoint - 1]) for row in got_sub}
    same = 0
    for row in ref_sub:
        mapped = tuple(-1 if v < 0 else id_map.get(int(v), -2) for v in row[:njoint - 1])
        same += mapped in got_rows
    return dict(ref_rows=len(ref_sub), got_rows=len(got_sub), identical_rows=same)


_report = {}


@pytest.mark.parametrize("fixture", sorted(BOUNDS))
def test_cuda_path_against_reference_run_with_real_networks(fixture):
    g = np.load(os.path.join(GOLD, fixture))
    mt = str(g["model_type"])
    H, W = int(g["h"]), int(g["w"])
    kw = {}
    if "init" in g.files:
        kw["init"] = str(g["init"])
    else:
        kw.update(gain=float(g["gain"]), head_gain=float(g["head_gain"]))
    torch.cuda.set_device(0)
    flat = O.make_flat_weights(mt, seed=int(g["weight_seed"]), **kw)
    body = isl_b200.Body(flat, mt, scale_search=g["scales"].tolist())
    cand, sub = body(synth.synth_frame(H, W, int(g["frame_seed"])))
    ref_c, ref_s = g["candidate"], g["subset"]
    # (1) maps: the float64 planes the peak kernels read, on the fixture's lattice
    st = int(g["lattice_stride"])
    ref_maps = g["heat_lattice"].astype(np.float64)
    got_maps = body._workspace(1, H, W)["heat"][0, :, ::st, ::st].cpu().numpy()
    assert got_maps.shape == ref_maps.shape
    map_scale = float(np.abs(ref_maps).max())
    map_err = float(np.abs(got_maps - ref_maps).max())
    assert ref_c.ndim == 2 and len(ref_c) > 50, "fixture without peaks proves nothing"
    stats, id_map = match_peaks(ref_c, cand, H, W, body.njoint)
    stats.update(compare_subsets(ref_s, sub, id_map, body.njoint))
    smax = float(np.abs(ref_c[:, 2]).max())
    stats["score_max_rel"] = stats["score_max"] / smax
    stats["map_max_abs"] = map_scale
    stats["map_err_max_rel"] = map_err / map_scale
    stats["map_err_mean_rel"] = float(np.abs(got_maps - ref_maps).mean()) / map_scale
    _report[fixture] = stats
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_drift.json"), "w") as f:
        json.dump(_report, f, indent=1, sort_keys=True)
    b = BOUNDS[fixture]
    n = stats["ref"]
    assert stats["map_err_max_rel"] <= MAP_TOL, stats
    assert stats["score_max_rel"] <= b["score"], stats
    if b["exact"] is not None:
        assert stats["exact"] >= b["exact"] * n, stats
        assert stats["exact"] + stats["near"] >= b["near"] * n, stats
        assert stats["missing"] <= b["unmatched"] * n and stats["extra"] <= b["unmatched"] * max(n, stats["got"]), stats
    # persons: random-init maps form few or none; where the reference found some, most must be found again
    if stats["ref_rows"] >= 10:
        assert abs(stats["got_rows"] - stats["ref_rows"]) <= 0.5 * stats["ref_rows"], stats
