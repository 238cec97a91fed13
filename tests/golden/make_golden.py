"""Generates the committed golden vectors by running the UNMODIFIED reference (/root/reference/src) on CPU.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs: tests/golden/*.npz. Every fixture stores the parameters needed to rebuild its inputs from seeds
(frames via isl_b200.synth.synth_frame, skeleton maps via isl_b200.synth, weights via
oracle.openpose_oracle.make_flat_weights) next to the outputs the reference produced for them, so the
fixtures stay a few KB each. Fixture kinds:
  net_<kind>.npz       reference nn.Module forward (model.py) on a small seeded input, weights loaded through
                       the reference's own util.transfer + load_state_dict path (body.py:35-36)
  body_<case>.npz      reference Body.__call__ + util.handDetect with a stub .model returning injected maps
  hand_<case>.npz      reference Hand.__call__ with a stub .model
  bodynet_<case>.npz   reference Body.__call__ end to end with the real network on a seeded frame
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import isl_b200  # noqa: E402,F401
from isl_b200 import synth  # noqa: E402
from oracle import openpose_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

BODY_CASES = [
    # name, model_type, H, W, people, skeleton seed, scale_search, dropped (person, joint) pairs
    ("coco_p3_s1", "coco", 240, 320, 3, 1, [0.5], []),
    ("coco_p2_s4", "coco", 310, 458, 2, 2, [0.5, 1.0, 1.5, 2.0], []),
    ("coco_p6_drop", "coco", 368, 496, 6, 3, [0.5, 1.0], [(0, 4), (1, 7), (2, 1), (3, 0), (3, 14)]),
    ("coco_p0_empty", "coco", 200, 264, 0, 4, [0.5], []),
    ("body25_p5_s2", "body25", 360, 480, 5, 1, [0.5, 1.0], []),
    ("body25_p12_s1", "body25", 480, 856, 12, 5, [1.0], [(0, 8), (4, 1), (7, 11)]),
    ("body25_p24_crowd", "body25", 540, 960, 24, 6, [0.5], []),
    ("body25_p1_odd", "body25", 187, 251, 1, 7, [0.5, 1.0, 1.5, 2.0], []),
]
HAND_CASES = [("hand_w109", 109, 3, [5]), ("hand_w150", 150, 4, []), ("hand_w64", 64, 5, [0, 20]),
              ("hand_w20", 20, 6, []), ("hand_w233_all_missing", 233, 7, list(range(21)))]
NET_CASES = [("coco", 40, 48), ("body25", 32, 56), ("hand", 48, 48)]


class Stub(torch.nn.Module):
    """Stands in for body_estimation.model / hand_estimation.model (body.py:63, hand.py:47)."""

    def __init__(self, fn):
        super().__init__()
        self.fn = fn

    def forward(self, data):
        outs = self.fn(data.numpy())
        if isinstance(outs, tuple):
            return tuple(torch.from_numpy(np.ascontiguousarray(o))[None] for o in outs)
        return torch.from_numpy(np.ascontiguousarray(outs))[None]


def body_stub_fn(model_type, skeletons, drop):
    def fn(data):
        return synth.render_maps(model_type, skeletons, data.shape[2] // 8, data.shape[3] // 8, drop=set(drop))

    return fn


def hand_points(seed, missing):
    pts = np.random.RandomState(seed).uniform(0.1, 0.9, (21, 2))
    for j in missing:
        pts[j] = -1
    return pts


reference_module = ref_import.reference_module


def main():
    _, _, _, rutil = ref_import.load()
    for (kind, h, w) in NET_CASES:
        flat = O.make_flat_weights(kind, seed=0)
        net = reference_module(kind, flat)
        x = torch.from_numpy(np.random.RandomState(11).uniform(-0.5, 0.5, (1, 3, h, w)).astype(np.float32))
        with torch.no_grad():
            out = net(x)
        outs = [out.numpy()] if kind == "hand" else [o.numpy() for o in out]
        np.savez_compressed(os.path.join(OUT, "net_%s.npz" % kind), h=h, w=w, input_seed=11, weight_seed=0,
                            **{"out%d" % i: o for i, o in enumerate(outs)})
        print("net", kind, [o.shape for o in outs], [float(np.abs(o).max()) for o in outs])

    for (name, mt, h, w, p, seed, scales, drop) in BODY_CASES:
        sk = synth.synth_skeletons(mt, p, seed)
        img = synth.synth_frame(h, w, seed)
        body = ref_import.make_body(mt, Stub(body_stub_fn(mt, sk, drop)), scales)
        cand, sub = body(img)
        boxes = rutil.handDetect(cand, sub, img) if len(sub) else []
        np.savez_compressed(os.path.join(OUT, "body_%s.npz" % name), model_type=mt, h=h, w=w, people=p, seed=seed,
                            scales=np.array(scales), drop=np.array(drop, dtype=np.int64).reshape(-1, 2),
                            candidate=cand, subset=sub,
                            boxes=np.array([[b[0], b[1], b[2], int(b[3])] for b in boxes], dtype=np.int64).reshape(-1, 4))
        print("body", name, cand.shape, sub.shape, len(boxes))

    for (name, w, seed, missing) in HAND_CASES:
        pts = hand_points(seed, missing)
        crop = synth.synth_frame(w, w, seed)
        hand = ref_import.make_hand(Stub(lambda d: synth.render_hand_maps(pts, d.shape[2] // 8, d.shape[3] // 8)))
        peaks = hand(crop)
        np.savez_compressed(os.path.join(OUT, "%s.npz" % name), w=w, seed=seed, missing=np.array(missing, dtype=np.int64),
                            peaks=peaks)
        print("hand", name, peaks[:3].tolist())

    # End to end with the real networks (random-init, gained so that the maps cross the thresholds).
    for (name, mt, h, w, scales, gain, head_gain) in [("coco_realnet", "coco", 240, 320, [0.5], 1.0, 1.0),
                                                      ("coco_realnet_gained", "coco", 184, 248, [1.0], 1.0, 3.0)]:
        flat = O.make_flat_weights(mt, seed=1, gain=gain, head_gain=head_gain)
        net = reference_module(mt, flat)
        img = synth.synth_frame(h, w, 21)
        body = ref_import.make_body(mt, net, scales)
        cand, sub = body(img)
        np.savez_compressed(os.path.join(OUT, "bodynet_%s.npz" % name), model_type=mt, h=h, w=w, frame_seed=21,
                            weight_seed=1, gain=gain, head_gain=head_gain, scales=np.array(scales),
                            candidate=cand, subset=sub)
        print("bodynet", name, cand.shape, sub.shape)


if __name__ == "__main__":
    main()
