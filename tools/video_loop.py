"""The frame loop from disk (N1): writes a synthetic MJPEG clip, then runs frames.VideoExtractor over it - every rank its
contiguous block of frames - and reports decode rate, end-to-end frames/s and which of the two limits the loop.

    python tools/video_loop.py [--frames 240] [--size 1280x720] [--model body25] [--batch 8]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/video_loop.py --frames 1920

One JSON line on rank 0. The clip is low-entropy (an up-scaled random pattern that scrolls), so the MJPEG decode cost per
frame is that of ordinary video, not of white noise; the networks run on whatever the decoder returns.
"""
import argparse
import json
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=240)
    ap.add_argument("--size", default="1280x720")
    ap.add_argument("--model", default="body25", choices=["coco", "body25"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--raw", action="store_true", help="a raw .npy dump instead of MJPEG (no decode cost)")
    ap.add_argument("--hands", type=int, default=2, help="fixed 128-px hand boxes per frame (random-init maps hold no persons "
                                                         "for util.handDetect); 0 = use handDetect")
    args = ap.parse_args()
    W, H = [int(v) for v in args.size.split("x")]

    import cv2
    import torch
    import torch.distributed as dist

    import isl_b200
    isl_b200.configure()
    from isl_b200 import frames as FR
    from isl_b200 import synth
    from isl_b200.extract import KeypointExtractor

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    work = os.path.join(tempfile.gettempdir(), "islpose_video_loop")
    path = os.path.join(work, "clip.npy" if args.raw else "clip.avi")
    if rank == 0:
        shutil.rmtree(work, ignore_errors=True)
        os.makedirs(work)
        base = cv2.resize(np.random.RandomState(0).randint(0, 256, (H // 16, W // 16, 3)).astype(np.uint8), (W, H),
                          interpolation=cv2.INTER_CUBIC)
        if args.raw:
            np.save(path, np.stack([np.roll(base, 7 * i, axis=1) for i in range(args.frames)]))
        else:
            wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30, (W, H))
            for i in range(args.frames):
                wr.write(np.roll(base, 7 * i, axis=1))
            wr.release()
    if world > 1:
        dist.barrier()
    body = isl_b200.Body(synth.make_flat_weights(args.model, seed=0), args.model, scale_search=[0.5, 1.0, 1.5, 2.0])
    hand = isl_b200.Hand(synth.make_flat_weights("hand", seed=0))
    ex = KeypointExtractor(body, hand)
    # decode alone: this rank's block, frames discarded
    t0 = time.perf_counter()
    n_dec = sum(len(idxs) for idxs, _ in FR.FrameFeeder(path, args.batch, rank, world))
    decode_alone = n_dec / (time.perf_counter() - t0)
    rates = []
    for rep in range(2):   # the first pass builds plans and buffers
        out = os.path.join(work, "out_%d_%d" % (rep, rank))
        boxes = [[(W // (args.hands + 1)) * (k + 1) - 64, H // 2 - 64, 128, k % 2 == 0] for k in range(args.hands)] or None
        vx = FR.VideoExtractor(ex, out, batch_size=args.batch, rank=rank, world_size=world, hand_boxes=boxes)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        rows = vx.extract_features_worker(path, "t", "e")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        rates.append((len(rows), dt, vx.stats))
    n, dt, stats = rates[-1]
    t = torch.tensor([dt, float(n), decode_alone], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dt, n, decode_alone = float(tmax[0]), float(tsum[1]), float(tsum[2])
    if rank == 0:
        fps = n / dt
        line = {"tool": "video_loop", "n_gpus": world, "frames": int(n), "size": args.size, "model": args.model,
                "source": "raw npy" if args.raw else "MJPEG avi via cv2.VideoCapture (%s)" % cv2.__version__,
                "frames_per_s": fps, "decode_alone_frames_per_s": decode_alone,
                "rank0_seconds": {k: round(stats[k], 3) for k in ("seconds", "decode_seconds", "rows_seconds",
                                                                      "writer_tail_seconds", "pipeline_wait_seconds")},
                "limiter": max((("decode (reader thread)", stats["decode_seconds"]),
                                ("host rows + JSON (Python; ~%d candidates per frame on random-init maps)" % (
                                    len(rows[0]["candidate"]) if rows else 0), stats["rows_seconds"] + stats["writer_tail_seconds"]),
                                ("GPU pipeline", stats["pipeline_wait_seconds"])), key=lambda kv: kv[1])[0],
                "hand_boxes_per_frame": args.hands,
                "note": "one reader thread per rank into pinned buffers; JSON per frame written by a writer thread"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
