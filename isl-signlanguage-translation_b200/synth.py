"""Synthetic inputs for tests and benchmarks: seeded frames, and injected stride-8 heat/PAF maps of P skeletons.

Random-init networks never produce a person (default init gives no peaks, gained init gives
hundreds of peaks and no persons), so everything downstream of the nets - peak NMS, PAF line integrals, greedy
matching, person assembly, handDetect - is exercised by handing both the reference `Body` and this package's
post-processing chain the *same* crafted network outputs. The reference only ever calls `self.model(data)`
(body.py:63, hand.py:47), so a stub module is a legal fixture on both sides.
"""
import numpy as np

from .tables import LIMB_SEQ, MAP_IDX, model_dims


def synth_frame(h, w, seed):
    """uint8 BGR frame exactly as SURVEY.md section 8d defines the benchmark inputs."""
    return np.random.RandomState(seed).randint(0, 256, (h, w, 3)).astype(np.uint8)


def synth_skeletons(model_type, n_people, seed, margin=0.08):
    """Joint positions in normalised image coordinates, shape [P, njoint-1, 2] (x, y) in (0, 1).

    Each skeleton is grown along limbSeq from the neck (joint 1) with limb lengths of 4-11 % of the frame, so
    every limb is far shorter than H/2 and the distance prior of body.py:158 stays at 0.
    """
    njoint, _ = model_dims(model_type)
    rng = np.random.RandomState(seed)
    limbs = LIMB_SEQ[model_type]
    out = np.zeros((n_people, njoint - 1, 2))
    for p in range(n_people):
        placed = {}
        placed[1] = rng.uniform(0.25, 0.75, 2)
        pending = list(limbs)
        while pending:
            rest = []
            for a, b in pending:
                if a in placed and b in placed:
                    continue
                if a in placed or b in placed:
                    src, dst = (a, b) if a in placed else (b, a)
                    ang = rng.uniform(0, 2 * np.pi)
                    ln = rng.uniform(0.04, 0.11)
                    pos = placed[src] + ln * np.array([np.cos(ang), np.sin(ang)])
                    placed[dst] = np.clip(pos, margin, 1 - margin)
                else:
                    rest.append((a, b))
            if len(rest) == len(pending):
                break
            pending = rest
        for j in range(njoint - 1):
            out[p, j] = placed.get(j, rng.uniform(margin, 1 - margin, 2))
    return out


def render_maps(model_type, skeletons, gh, gw, sigma=1.0, peak=0.9, drop=None):
    """Stride-8 network outputs for the given skeletons: (paf [npaf, gh, gw], heat [njoint, gh, gw]) float32.

    heat[j] = max over people of peak * exp(-d^2 / (2 sigma^2)) (d in grid cells), last channel = 1 - max;
    PAF channels of limb k hold the unit limb vector within one cell of the segment.
    `drop` is an optional set of (person, joint) pairs to leave out (missing-part cases).
    """
    njoint, npaf = model_dims(model_type)
    ys, xs = np.mgrid[0:gh, 0:gw].astype(np.float64)
    heat = np.zeros((njoint, gh, gw), np.float64)
    paf = np.zeros((npaf, gh, gw), np.float64)
    drop = drop or set()
    for p in range(skeletons.shape[0]):
        for j in range(njoint - 1):
            if (p, j) in drop:
                continue
            cx = skeletons[p, j, 0] * gw - 0.5
            cy = skeletons[p, j, 1] * gh - 0.5
            g = peak * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * sigma * sigma))
            heat[j] = np.maximum(heat[j], g)
        for k, (a, b) in enumerate(LIMB_SEQ[model_type]):
            if (p, a) in drop or (p, b) in drop:
                continue
            ax, ay = skeletons[p, a, 0] * gw - 0.5, skeletons[p, a, 1] * gh - 0.5
            bx, by = skeletons[p, b, 0] * gw - 0.5, skeletons[p, b, 1] * gh - 0.5
            vx, vy = bx - ax, by - ay
            ln = max(np.hypot(vx, vy), 1e-6)
            ux, uy = vx / ln, vy / ln
            t = np.clip(((xs - ax) * ux + (ys - ay) * uy), 0, ln)
            d = np.hypot(xs - (ax + t * ux), ys - (ay + t * uy))
            m = d <= 1.0
            cxi, cyi = MAP_IDX[model_type][k]
            paf[cxi][m] = ux
            paf[cyi][m] = uy
    heat[njoint - 1] = 1.0 - heat[:njoint - 1].max(0)
    return paf.astype(np.float32), heat.astype(np.float32)


def render_hand_maps(points, gh, gw, sigma=1.0, peak=0.9):
    """Stride-8 hand network output [22, gh, gw] for 21 key points given in normalised crop coordinates
    (entries with a negative coordinate are absent)."""
    ys, xs = np.mgrid[0:gh, 0:gw].astype(np.float64)
    heat = np.zeros((22, gh, gw), np.float64)
    for j in range(21):
        if points[j, 0] < 0:
            continue
        cx = points[j, 0] * gw - 0.5
        cy = points[j, 1] * gh - 0.5
        heat[j] = peak * np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * sigma * sigma))
    heat[21] = 1.0 - heat[:21].max(0)
    return heat.astype(np.float32)


def make_flat_weights(kind, seed=0, init="torch", gain=1.0, head_gain=1.0):
    """Seeded random weights in the reference's on-disk format - a flat dict of Caffe layer names
    ('conv1_1.weight', 'Mprelu1_stage0_L2_0.weight', ...) -> float32 tensors (what torch.load returns for the shipped
    .pth files, body.py:35-36). No trained weights ship with the reference, so benchmarks and tools run on these.
      init="torch": the distribution nn.Conv2d / nn.PReLU are constructed with (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for
                    weight and bias, PReLU slope 0.25): SURVEY.md section 8d's "random-init weights"
      init="he":    He-uniform scaled by `gain` (activations stay alive, thousands of peaks: a stress input)
    Layers come from the launch program (nets.build_program), in network order."""
    import math

    import torch

    from .nets import build_program

    g = torch.Generator().manual_seed(seed)
    w = {}
    for step in build_program(kind).steps:
        if step[0] != "conv" or step[1]["layer"] + ".weight" in w:
            continue
        s = step[1]
        name, cout, k = s["layer"], s["cout"], s["k"]
        cin = 3 if s["first"] else (sum(1 for c in s["chan_map"] if c is not None) if s["chan_map"] else s["src"][2])
        if init == "torch":
            bound = bias_bound = 1.0 / math.sqrt(cin * k * k)
        else:
            bound, bias_bound = gain * math.sqrt(6.0 / (cin * k * k)), 0.05
            if name.startswith("Mconv7") or name in ("conv5_5_CPM_L1", "conv5_5_CPM_L2", "conv6_2_CPM"):
                bound *= head_gain
        w[name + ".weight"] = (torch.rand((cout, cin, k, k), generator=g) * 2 - 1) * bound
        w[name + ".bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bias_bound
        if s["prelu"] is not None:
            w[s["prelu"] + ".weight"] = torch.full((cout,), 0.25) if init == "torch" else torch.rand((cout,), generator=g) * 0.3
    return w
