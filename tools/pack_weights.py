"""Converts a reference weight file (torch.save archive or .caffemodel) into this package's packed blob, once:

    python tools/pack_weights.py model/body_pose_model.pth model/body_pose_model.islpose

Body(path) / Hand(path) accept either file; the blob holds the conv weights as the bf16 values the kernels consume
(half the size, identical results) and is read with one np.fromfile. No torch, no GPU."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import isl_b200  # noqa: E402,F401
from isl_b200 import weights  # noqa: E402

if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    flat = weights.load_flat(sys.argv[1])
    weights.write_packed(sys.argv[2], flat)
    print("%d tensors, %.1f MB -> %.1f MB" % (len(flat), os.path.getsize(sys.argv[1]) / 1e6, os.path.getsize(sys.argv[2]) / 1e6))
