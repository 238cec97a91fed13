#!/usr/bin/env python
"""Benchmark of the OpenPose keypoint-extraction hot path (body + hand) - one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W [--workload C2|C3] [--batch B] [--impl reference]

A step = one pass of body + hand extraction over one batch of synthetic frames per rank (frames shard across ranks,
no collective on the data path; scaling is weak). Workloads (BASELINE.json configs):
  C2  coco body + hand, 640x480, scale_search [0.5,1,1.5,2], two fixed hand boxes per frame      (default)
  C3  body25 + hand, 1280x720, same scales, two 128-px hand boxes per frame (C4's clips are batches of C3 frames)
  C5  body25 + hand, 1920x1080, same scales, 40 hand boxes per frame (the multi-person stress shape)
Hand boxes are fixed per workload because random-init weights never produce a person for util.handDetect.

value   frames/s with the frames already resident in HBM (device-timed, max over ranks)
e2e     frames/s through the public host API (numpy frames in, numpy results out; H2D from pinned memory and the
        D2H of candidate / subset / hand peaks inside the timed region)
roofline      the tcgen05 conv kernels: algorithmic FLOPs of a step's network replays / the CUDA-event time of those
              replays run on their own (a step overlaps them with the post-processing, so they cannot be timed inside it)
post_roofline the body map post-processing (x8 cubic, resize to frame size, scale accumulation, gaussian, NMS) of one
              chunk run on its own: SURVEY.md 8d bytes / CUDA-event time against the measured HBM copy bandwidth
cpu_baseline  the oracle (restated reference, calling cv2 / scipy / torch-CPU where the reference does) on the host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec body+hand keypoints @368 4-scale"
SCALES = [0.5, 1.0, 1.5, 2.0]
WEIGHT_INIT = "torch"   # nn.Conv2d's default init distribution; "he" = noisy-map stress
WORKLOADS = {
    # name: (model_type, H, W, hand boxes [x, y, w, is_left], default batch per rank)
    "C2": ("coco", 480, 640, [[400, 250, 109, True], [22, 246, 90, False]], 16),
    "C3": ("body25", 720, 1280, [[800, 300, 128, True], [300, 300, 128, False]], 16),
    # C5 (BASELINE.json configs[4]): 1080p, "20+ people" = 40 hand crops per frame on a fixed 8 x 5 lattice, 96..127 px
    "C5": ("body25", 1080, 1920, [[40 + 230 * (i % 8), 60 + 200 * (i // 8), 96 + (7 * i) % 32, i % 2 == 0] for i in range(40)], 4),
}


def workload_name(wl, batch):
    mt, H, W, boxes, _ = WORKLOADS[wl]
    sizes = ",".join("%dpx" % b[2] for b in boxes) if len(boxes) <= 4 else "%d..%dpx" % (min(b[2] for b in boxes), max(b[2] for b in boxes))
    return "%s: %s body + hand, %dx%d, scale_search %s, %d hand boxes/frame (%s), batch %d frames/step/rank" % (
        wl, mt, W, H, SCALES, len(boxes), sizes, batch)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        p = json.load(open(path))
        return p.get("bf16_tflops_sustained", 1379.5), p.get("bf16_tflops", 1636.0), p.get("hbm_gbs", 6542.7), "measured"
    return 1400.0, 1590.0, 6500.0, "fallback"


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant conv kernel from the committed
# `ncu --set full` capture (profiles/): filled in by hand from the capture, None until one exists for this round
CONV_DRAM_TRAFFIC = 39.79e6   # v5, 7x7 128->128, 92x164x8: 36.10 MB read + 3.69 MB written (profiles/r1_ncu_full_conv7x7_v5.txt)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.rows = []
        self.stop_flag = False
        self.index = index

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference_frame(wl, seed, nets=None):
    """One frame of the workload through the oracle port of the reference (backend 'lib': cv2 / scipy / torch CPU
    exactly where the reference calls them). Returns (seconds, nets) so repeated calls reuse the weights."""
    import torch

    from isl_b200 import synth
    from oracle import openpose_oracle as O

    mt, H, W, boxes, _ = WORKLOADS[wl]
    if nets is None:
        torch.set_num_threads(os.cpu_count() or 1)
        nets = (O.make_net_fn(mt, synth.make_flat_weights(mt, seed=0, init=WEIGHT_INIT)),
                O.make_net_fn("hand", synth.make_flat_weights("hand", seed=0, init=WEIGHT_INIT)))
    frame = synth.synth_frame(H, W, seed)
    t0 = time.perf_counter()
    try:
        O.body_call(nets[0], frame, mt, tuple(SCALES), backend="lib", strict=False)
    except IndexError:
        pass
    for (x, y, w, _) in boxes:
        O.hand_call(nets[1], np.ascontiguousarray(frame[y:y + w, x:x + w, :]), backend="lib")
    return time.perf_counter() - t0, nets


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the Python reference cannot
    travel to the GPU box) on all host threads, one frame of the workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    budget_s = 240.0
    warm, nets = cpu_reference_frame(wl, 999)
    times = []
    spent = warm
    for w in range(max(args.warmup - 1, 0)):
        if spent + warm > budget_s * 0.4:
            break
        t, nets = cpu_reference_frame(wl, 998 - w, nets)
        spent += t
    for k in range(args.steps):
        if times and spent + max(times) > budget_s:
            break
        t, nets = cpu_reference_frame(wl, k, nets)
        times.append(t)
        spent += t
    fps = len(times) / sum(times)
    cores = os.cpu_count() or 1
    sample = "1 frame of the workload per step on %d host threads; %d of %d requested steps timed inside a %.0f s budget" % (
        cores, len(times), args.steps, budget_s)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": len(times),
            "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(wl, 1), "timing": "host wall clock around each step (CPU only, no device work)"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="frames per step per rank (0 = workload default)")
    ap.add_argument("--chunk", type=int, default=0, help="split a step into pipeline stages of this many frames (0 = one stage per step)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--weights", default="torch", choices=["torch", "he"], help="random-init distribution (he = noisy-map stress)")
    args = ap.parse_args()
    global WEIGHT_INIT
    WEIGHT_INIT = args.weights
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import isl_b200
    from isl_b200 import _lib, synth
    from isl_b200.extract import KeypointExtractor

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    mt, H, W, boxes, default_batch = WORKLOADS[args.workload]
    B = args.batch or default_batch
    body = isl_b200.Body(synth.make_flat_weights(mt, seed=0, init=WEIGHT_INIT), mt, scale_search=SCALES)
    hand = isl_b200.Hand(synth.make_flat_weights("hand", seed=0, init=WEIGHT_INIT))
    ex = KeypointExtractor(body, hand, chunk=args.chunk or None)
    hand_boxes = [boxes] * B

    def frames_for(step):
        return [synth.synth_frame(H, W, (rank * 100003 + step * B + i) % (2 ** 31)) for i in range(B)]

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")  # > 126 MB of L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L = _lib.lib()

    # ---- leg 1: device-resident inputs (value) ---------------------------------------------------------------
    total_steps = args.warmup + args.steps
    dev_frames = [torch.from_numpy(np.stack(frames_for(s))).cuda() for s in range(min(total_steps, 4))]

    def run_steps(batches_of, n_steps):
        """K steps through the two-lane pipeline; returns (device ms between the brackets, results of the last step)."""
        def feed():
            for s in range(n_steps):
                flush.zero_()   # evicts L2 between steps (the per-step working set is far larger than L2 anyway)
                yield batches_of(s), hand_boxes
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        last = None
        for res in ex.pipeline(feed()):
            last = res
        e1.record()
        barrier()
        return e0.elapsed_time(e1), last

    run_steps(lambda s: dev_frames[s % len(dev_frames)], args.warmup)   # W untimed steps (both lanes get their buffers)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.islpose_launch_count()
    dev_ms, _ = run_steps(lambda s: dev_frames[(args.warmup + s) % len(dev_frames)], args.steps)
    gpu_launches = L.islpose_launch_count() - launches0
    sampler.stop_flag = True

    # ---- leg 2: host API end to end (e2e) ------------------------------------------------------------------
    host_frames = [frames_for(1000 + s) for s in range(min(args.steps, 4))]
    run_steps(lambda s: host_frames[s % len(host_frames)], 2)
    e2e_ms, res = run_steps(lambda s: host_frames[s % len(host_frames)], args.steps)
    d2h = sum(c.nbytes + sb.nbytes + sum(p.nbytes for p in hp) for c, sb, hp in res)
    h2d = B * H * W * 3   # hand crops are cut from the device copy of the frame

    # ---- leg 3: the convolution plans of one step on their own (roofline of the tcgen05 kernels) -------------
    # exactly the network replays a step performs (same chunks, lanes and streams; resize, im2col and max-pool
    # launches included), without the post-processing, so the events time tensor-bound work only
    chunk = ex.chunk or B
    lanes = ex._lane_streams(torch)
    main = torch.cuda.current_stream()

    def networks_only(frames_dev):
        flops = 0
        ready = torch.cuda.Event()
        ready.record(main)
        for ci, a in enumerate(range(0, B, chunk)):
            sub = frames_dev[a:a + chunk]
            with torch.cuda.stream(lanes[ci % 2]):
                lanes[ci % 2].wait_event(ready)
                body.model.timing = []
                body.network_outputs(sub, H, W, lane=ci % 2)
                flops += sum(t[2] for t in body.model.timing)
            with torch.cuda.stream(lanes[2 + ci % 2]):
                lanes[2 + ci % 2].wait_event(ready)
                crops = [frames_dev[a + i, y:y + w, x:x + w, :].contiguous() for i in range(sub.shape[0]) for (x, y, w, _) in boxes]
                hand.model.timing = []
                for ca in range(0, len(crops), hand.MAX_CROPS_PER_REPLAY):   # as Hand.enqueue replays them
                    hand.network_outputs(crops[ca:ca + hand.MAX_CROPS_PER_REPLAY], lane=ci % 2)
                flops += sum(t[2] for t in hand.model.timing)
        for st in lanes:
            main.wait_stream(st)
        body.model.timing = hand.model.timing = None
        return flops

    networks_only(dev_frames[0])
    barrier()
    conv_events, conv_flops = [], 0
    for s in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        conv_flops += networks_only(dev_frames[s % len(dev_frames)])
        e1.record()
        conv_events.append((e0, e1))
    barrier()
    conv_ms = sum(a.elapsed_time(b) for a, b in conv_events)

    # ---- leg 4: the body map post-processing on its own (HBM roofline of subsystem 2) -------------------------
    # maps accumulation + gaussian/NMS/peak lists for one chunk, from the network outputs left by leg 3
    nb = min(chunk, B)
    ws = body._workspace(nb, H, W, 0)
    maps = body.network_outputs(dev_frames[0][:nb], H, W, lane=0)
    parts = body.njoint - 1
    heat_scales = body._scales_struct(maps, 1)
    need = L.islpose_maps_workspace_floats(heat_scales, len(maps), nb, parts)
    if ws.get("mid") is None or ws["mid"].numel() < need:
        ws["mid"] = torch.empty((need,), dtype=torch.float32, device="cuda")

    def post_maps():
        _lib.check(L.islpose_maps_accumulate(heat_scales, len(maps), body.njoint, nb, H, W, parts, 1, _lib.ptr(ws["heat"]),
                                             _lib.ptr(ws["mid"]), ws["mid"].numel(), _lib.stream_ptr()), "maps_accumulate")
        _lib.check(L.islpose_body_peaks(_lib.ptr(ws["heat"]), nb * parts, H, W, body._gauss, body.thre1, 1024,
                                        _lib.ptr(ws["counts"]), _lib.ptr(ws["keys"]), _lib.ptr(ws["scores"]),
                                        _lib.ptr(ws["overflow"]), _lib.stream_ptr()), "body_peaks")

    post_maps()
    barrier()
    post_events = []
    for s in range(args.steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        post_maps()
        e1.record()
        post_events.append((e0, e1))
    barrier()
    post_ms = sum(a.elapsed_time(b) for a, b in post_events) / args.steps
    ws["overflow"].zero_()
    # SURVEY.md section 8d convention for the bytes of the reference dataflow, heat maps only (the PAF maps are never
    # materialised here; their share of B_post is reported separately as what the lazy sampler avoids)
    S = len(SCALES)
    grid_bytes = sum(4 * (m[2][2] // 8) * (m[2][3] // 8) for m in maps)
    post_bytes_heat = nb * (4 * H * W * ((2 * S - 1) * body.njoint + 2 * parts) + grid_bytes * body.njoint)
    post_bytes_all = nb * (4 * H * W * ((2 * S - 1) * (body.njoint + body.npaf) + 2 * parts) + grid_bytes * (body.njoint + body.npaf))

    # ---- single-frame latency through the same public API (BASELINE.json's C2 is literally one frame) -------
    one = [host_frames[0][0]]
    for _ in range(2):
        ex.batch(one, hand_boxes[:1])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        ex.batch(one, hand_boxes[:1])
    torch.cuda.synchronize()
    single_ms = (time.perf_counter() - t0) / 3 * 1e3

    times = torch.tensor([dev_ms, e2e_ms, conv_ms, post_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, conv_ms, post_ms = [float(x) for x in times.cpu()]

    if rank == 0:
        sustained, burst, hbm_peak, src = peaks()
        achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
        frames_total = world * B * args.steps
        line = {
            "metric": METRIC, "value": frames_total / (dev_ms * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args.workload, B), "weights": "seeded random init, %s (no trained weights ship with the reference)" % (
                           "nn.Conv2d default distribution" if WEIGHT_INIT == "torch" else "He-uniform (noisy maps: thousands of peaks)"),
                       "l2": "flushed with a 256 MiB write before every timed step",
                       "timing": "CUDA events on the launching stream around the K steps; max over ranks",
                       "pipeline": "steps run through KeypointExtractor.pipeline(): two lanes, the post-processing and copies of "
                                   "one step overlap the convolutions of the next; all K steps start and end inside the timed region",
                       "single_frame_latency_ms": round(single_ms, 2)},
            "clocks": sampler.summary(),
            "e2e": {"value": frames_total / (e2e_ms * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(gpu_launches),
            "roofline": {"bound": "tensor", "kernel": "conv_umma_* (tcgen05 implicit-GEMM variants): the network replays of "
                                                      "a step run on their own, same chunks / lanes / streams as the step; the "
                                                      "resize, first-layer and max-pool launches are inside the same events",
                         "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                         "peak_source": "%s bf16_tflops_sustained (burst %.1f)" % (src, burst),
                         "traffic": CONV_DRAM_TRAFFIC, "conv_share_of_step": conv_ms / max(dev_ms, 1e-9)},
            "post_roofline": {"bound": "hbm", "kernel": "upsample8 + resize_accumulate + gauss_window + sort_peaks "
                                                        "(body maps of one %d-frame chunk, run on their own)" % nb,
                              "achieved": post_bytes_heat / (post_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                              "frac": post_bytes_heat / (post_ms * 1e-3) / 1e9 / hbm_peak,
                              "limiter": "instruction issue and the FP64 pipe, not HBM: the reference's float32 (no FMA, fixed order) "
                                         "cubic stages and float64 accumulation / gaussian are reproduced bit for bit "
                                         "(profiles/r1_ncu_full_post.txt, DESIGN.md section 7)",
                              "ms_per_frame": post_ms / nb, "algorithmic_bytes_per_frame": post_bytes_heat // nb,
                              "reference_dataflow_bytes_per_frame_incl_paf": post_bytes_all // nb},
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample of the same workload on the host cores: one untimed frame (weights, thread pools), then
            # frames until about 12 s of CPU work have been timed (at least 2, at most 4)
            _, nets = cpu_reference_frame(args.workload, 0)
            ts = []
            while len(ts) < 2 or (sum(ts) < 12.0 and len(ts) < 4):
                t, nets = cpu_reference_frame(args.workload, 1 + len(ts), nets)
                ts.append(t)
            line["cpu_baseline"] = {"value": len(ts) / sum(ts), "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": "%d frames of the workload (body 4 scales + %d hands each) after one untimed frame, "
                                              "%.1f s of CPU work" % (len(ts), len(boxes), sum(ts))}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
