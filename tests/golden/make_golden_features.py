"""Golden vectors for the feature-vector step, produced by the UNMODIFIED reference functions:
util.get_bodypose / util.get_handpose (imported from /root/reference/src through oracle.ref_import) and
populate_features (ISL_model_xy.py:78-112; that file runs model loading at import, so the function's own source text
is extracted with ast and executed as is).

Run in the build container only:   python tests/golden/make_golden_features.py
Inputs: candidate / subset of the committed body fixtures + seeded hand key points. Output: tests/golden/features.npz
"""
import ast
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def reference_populate_features():
    path = "/root/reference/ISL_model_xy.py"
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "populate_features"][0]
    ns = {"np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns["populate_features"]


def hand_peaks(seed, n_hands):
    rs = np.random.RandomState(seed)
    hands = []
    for _ in range(n_hands):
        p = rs.randint(1, 900, (21, 2)).astype(np.int64)
        p[rs.rand(21) < 0.3] = 0   # [0, 0] = key point not found (hand.py:66)
        hands.append(p)
    return hands


def main():
    _, _, _, ref_util = ref_import.load()
    populate = reference_populate_features()
    out = {}
    cases = sorted(glob.glob(os.path.join(OUT, "body_*.npz")))
    for k, path in enumerate(cases):
        d = np.load(path, allow_pickle=True)
        name = os.path.basename(path)[5:-4]
        mt = str(d["model_type"])
        cand, sub = d["candidate"], d["subset"]
        n_hands = k % 3
        hands = hand_peaks(100 + k, n_hands)
        circles, sticks = ref_util.get_bodypose(cand, sub, mt)
        edges, peaks = ref_util.get_handpose(hands)
        feat = populate(circles, peaks)
        out[name + "/hands"] = np.stack(hands) if hands else np.zeros((0, 21, 2), dtype=np.int64)
        out[name + "/circles"] = np.array(circles, dtype=np.float64).reshape(-1, 2)
        out[name + "/sticks"] = np.array(sticks, dtype=np.float64).reshape(-1, 4)
        out[name + "/n_edges"] = np.array([len(e) for e in edges])
        out[name + "/feature"] = np.asarray(feat, dtype=np.float64)
        print(name, mt, "circles", len(circles), "sticks", len(sticks), "hands", n_hands, "feature", feat.shape)
    np.savez_compressed(os.path.join(OUT, "features.npz"), **out)


if __name__ == "__main__":
    main()
