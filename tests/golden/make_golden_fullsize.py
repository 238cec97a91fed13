"""Full-size golden vectors from the UNMODIFIED reference (/root/reference/src), BASELINE.json's shapes.

Run in the build container only (needs /root/reference; minutes of CPU):   python tests/golden/make_golden_fullsize.py [case ...]

  bodynet_c2_coco_s4.npz        configs[1] exactly: reference Body('coco').__call__ with the real (seeded, nn.Conv2d-default
                                init) network on the 640x480 frame of seed 0, scale_search [0.5, 1, 1.5, 2]
  bodynet_c3_body25_s4.npz      configs[2]'s frame shape: reference Body('body25') with the real He-initialised network
                                (default init never crosses thre1 on body25) on a 1280x720 frame, four scales
  body_body25_p24_720p_s4.npz   injected maps (isl_b200.synth), 24 people at 1280x720, four scales: candidate AND subset
  body_body25_p40_1080p_s4.npz  injected maps, 40 people at 1920x1080, four scales (configs[4], SURVEY.md section 8d: P = 40)

The real-network fixtures hold what the reference computes in fp32 on this container's cv2 (IPP resize) / scipy / torch;
the CUDA path computes its networks in bf16, so tests/test_gpu_reference_e2e.py compares with stated drift bounds.
The injected-map fixtures are compared bit for bit.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import isl_b200  # noqa: E402,F401
from isl_b200 import synth  # noqa: E402
from oracle import openpose_oracle as O  # noqa: E402
from oracle import ref_import  # noqa: E402
from make_golden import Stub, body_stub_fn, reference_module  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SCALES = [0.5, 1.0, 1.5, 2.0]

REALNET = {
    # name: (model_type, H, W, frame seed, weight seed, init)
    "bodynet_c2_coco_s4": ("coco", 480, 640, 0, 0, "torch"),
    "bodynet_c3_body25_s4": ("body25", 720, 1280, 500, 0, "he"),
}
INJECTED = {
    # name: (model_type, H, W, people, skeleton seed)
    "body_body25_p24_720p_s4": ("body25", 720, 1280, 24, 11),
    "body_body25_p40_1080p_s4": ("body25", 1080, 1920, 40, 12),
}


def main(argv):
    want = set(argv) or set(REALNET) | set(INJECTED)
    _, _, _, rutil = ref_import.load()
    torch.set_num_threads(os.cpu_count() or 1)
    for name, (mt, H, W, fseed, wseed, init) in REALNET.items():
        if name not in want:
            continue
        t0 = time.time()
        flat = O.make_flat_weights(mt, seed=wseed, init=init)
        body = ref_import.make_body(mt, reference_module(mt, flat), SCALES)
        img = synth.synth_frame(H, W, fseed)
        with torch.no_grad():
            cand, sub = body(img)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), model_type=mt, h=H, w=W, frame_seed=fseed, weight_seed=wseed,
                            init=init, scales=np.array(SCALES), candidate=cand, subset=sub)
        print(name, cand.shape, sub.shape, "%.0f s" % (time.time() - t0), flush=True)
    for name, (mt, H, W, people, seed) in INJECTED.items():
        if name not in want:
            continue
        t0 = time.time()
        sk = synth.synth_skeletons(mt, people, seed)
        img = synth.synth_frame(H, W, seed)
        body = ref_import.make_body(mt, Stub(body_stub_fn(mt, sk, [])), SCALES)
        cand, sub = body(img)
        boxes = rutil.handDetect(cand, sub, img) if len(sub) else []
        np.savez_compressed(os.path.join(OUT, name + ".npz"), model_type=mt, h=H, w=W, people=people, seed=seed,
                            scales=np.array(SCALES), drop=np.zeros((0, 2), dtype=np.int64), candidate=cand, subset=sub,
                            boxes=np.array([[b[0], b[1], b[2], int(b[3])] for b in boxes], dtype=np.int64).reshape(-1, 4))
        print(name, cand.shape, sub.shape, len(boxes), "%.0f s" % (time.time() - t0), flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
