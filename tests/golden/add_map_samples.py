"""Adds a lattice sample of the reference's float64 heatmap_avg (body.py:48,80) to the real-network fixtures.

Run in the build container after make_golden.py / make_golden_fullsize.py:   python tests/golden/add_map_samples.py

The reference does not return its maps, so they are recomputed with the oracle in `lib` mode (the very cv2 / scipy /
torch-CPU calls the reference makes) from the same seeds, and the script first asserts that the oracle's peak table for
those maps equals the fixture's reference-produced candidate table (positions and ids exactly, scores to 1e-6: torch-CPU
thread partitioning); only then are the samples stored. `heat_lattice` = heatmap_avg[::s, ::s, :njoint-1] as float32,
`lattice_stride` = s. tests/test_gpu_reference_e2e.py compares the CUDA path's maps against them within the bf16 tolerance.
"""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import isl_b200  # noqa: E402,F401
from isl_b200 import synth  # noqa: E402
from oracle import openpose_oracle as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    for path in sorted(glob.glob(os.path.join(OUT, "bodynet_*.npz"))):
        g = dict(np.load(path))
        mt = str(g["model_type"])
        H, W = int(g["h"]), int(g["w"])
        kw = {"init": str(g["init"])} if "init" in g else dict(gain=float(g["gain"]), head_gain=float(g["head_gain"]))
        flat = O.make_flat_weights(mt, seed=int(g["weight_seed"]), **kw)
        img = synth.synth_frame(H, W, int(g["frame_seed"]))
        heat_avg, _ = O.body_maps(O.make_net_fn(mt, flat), img, mt, tuple(g["scales"].tolist()), backend="lib")
        njoint = 26 if mt == "body25" else 19
        peaks = O.body_peaks(heat_avg, njoint, backend="lib")
        table = np.array([list(p) for part in peaks for p in part], dtype=np.float64).reshape(-1, 4)
        ref = g["candidate"]
        assert table.shape == ref.shape and np.array_equal(table[:, [0, 1, 3]], ref[:, [0, 1, 3]]), path
        assert np.abs(table[:, 2] - ref[:, 2]).max() <= 1e-6, path
        s = 32 if H * W > 600000 else 16
        g["heat_lattice"] = np.ascontiguousarray(np.transpose(heat_avg[::s, ::s, :njoint - 1], (2, 0, 1))).astype(np.float32)
        g["lattice_stride"] = np.int64(s)
        np.savez_compressed(path, **g)
        print(os.path.basename(path), g["heat_lattice"].shape, "max |heat| %.4g" % np.abs(g["heat_lattice"]).max(), flush=True)


if __name__ == "__main__":
    main()
