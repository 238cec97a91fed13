"""One launch of every non-convolution kernel of a body + hand call inside a cudaProfilerStart/Stop range, for
`ncu --profile-from-start off --set full` (tools/gpu_ncu.sh): resize, first layer, map accumulation, gaussian + NMS, peak
sort, PAF end points / scoring, matching, assembly, feature rows, hand heat maps / gaussian / selection.

    python tools/ncu_kernels.py [C2|C3] [frames]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import isl_b200  # noqa: E402
isl_b200.configure()
from isl_b200 import synth  # noqa: E402
from isl_b200.extract import KeypointExtractor  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    mt, H, W, boxes, _ = bench.WORKLOADS[wl]
    torch.cuda.set_device(0)
    # the networks run kernel by kernel here (graph=False) so that the profiler range holds plain launches
    body = isl_b200.Body(synth.make_flat_weights(mt, seed=0, init="torch"), mt, scale_search=bench.SCALES, tuning={"graph": False})
    hand = isl_b200.Hand(synth.make_flat_weights("hand", seed=0, init="torch"), tuning={"graph": False})
    ex = KeypointExtractor(body, hand)
    frames = [synth.synth_frame(H, W, i) for i in range(nb)]
    hb = [boxes] * nb
    ex.batch(frames, hb)          # warm-up: plans, workspaces, staging buffers
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    res = ex.batch(frames, hb)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("%s x%d: %d candidates in frame 0" % (wl, nb, len(res[0][0])))


if __name__ == "__main__":
    main()
