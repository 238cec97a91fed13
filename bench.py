#!/usr/bin/env python
"""Benchmark of the OpenPose keypoint-extraction hot path (body + hand) - one JSON line on stdout.

    python bench.py --gpus N --steps K --warmup W [--workload C2|C3|C4|C5] [--batch B] [--no-sub] [--impl reference]

A step = one pass of body + hand extraction over one batch of synthetic frames per rank (frames shard across ranks,
no collective on the data path; scaling is weak). Workloads (BASELINE.json configs):
  C2  coco body + hand, 640x480, scale_search [0.5,1,1.5,2], two fixed hand boxes per frame      (default)
  C3  body25 + hand, 1280x720, same scales, two 128-px hand boxes per frame, batch 32 (configs[2])
  C4  the frame loop of the reference's extract_features*.py on C3-shaped frames: 30-frame clips of host frames, sharded
      by frame index over the ranks, run through one pipeline like a video, per-frame feature rows formed by every rank
      (as the reference's workers keep their own CSV rows), the per-frame results gathered on rank 0 and merged in frame
      order inside the timed region (configs[3]); one clip per rank per step
  C5  body25 + hand, 1920x1080, same scales, 40 hand boxes per frame (the multi-person stress shape, configs[4])
Hand boxes are fixed per workload because random-init weights never produce a person for util.handDetect.
The JSON line is the headline workload (C2 unless --workload says otherwise); the other workloads are measured in the
same run with fewer steps and reported under "sub_results", each with its own value / e2e / roofline.

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
    "C3": ("body25", 720, 1280, [[800, 300, 128, True], [300, 300, 128, False]], 32),
    "C4": ("body25", 720, 1280, [[800, 300, 128, True], [300, 300, 128, False]], 30),
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


def conv_dram_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant conv kernel, read from the committed
    `ncu --set full` capture of this round (profiles/r2_ncu_conv_dram.csv: `ncu -i ... --page raw --csv` of the v5 kernel
    on the 7x7 128->128 layer). None when no capture is committed."""
    import csv

    path = os.path.join(ROOT, "profiles", "r2_ncu_conv_dram.csv")
    if not os.path.isfile(path):
        return None, None
    try:
        rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
        hdr = rows[0]
        units = rows[1]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        total = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            vals = [float(r[i].replace(",", "")) * scale[units[i]] for r in rows[2:] if len(r) > i and r[i]]
            total += sum(vals) / len(vals)
        return total, os.path.relpath(path, ROOT)
    except Exception:
        return None, None



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
    """One frame of the workload on the host cores. With the reference's own files at hand (oracle/_ref/, placed there by
    __graft_entry__.build() in the build container; /root/reference itself there) this runs the UNMODIFIED reference:
    its Body.__call__ / Hand.__call__ over its nn.Modules (kind "reference"). Otherwise the oracle port in `lib` mode:
    cv2 / scipy / torch CPU exactly where the reference calls them (kind "port"). Returns (seconds, state)."""
    import torch

    from isl_b200 import synth
    from oracle import openpose_oracle as O
    from oracle import ref_import

    mt, H, W, boxes, _ = WORKLOADS[wl]
    if nets is None:
        torch.set_num_threads(os.cpu_count() or 1)
        flat_b = synth.make_flat_weights(mt, seed=0, init=WEIGHT_INIT)
        flat_h = synth.make_flat_weights("hand", seed=0, init=WEIGHT_INIT)
        if ref_import.available():
            nets = ("reference", ref_import.make_body(mt, ref_import.reference_module(mt, flat_b), SCALES),
                    ref_import.make_hand(ref_import.reference_module("hand", flat_h)))
        else:
            nets = ("port", O.make_net_fn(mt, flat_b), O.make_net_fn("hand", flat_h))
    frame = synth.synth_frame(H, W, seed)
    t0 = time.perf_counter()
    with torch.no_grad():
        try:
            if nets[0] == "reference":
                nets[1](frame)
            else:
                O.body_call(nets[1], frame, mt, tuple(SCALES), backend="lib", strict=False)
        except IndexError:
            pass   # body.py:193-197 raises on a third matching row (noisy maps); the frame's time still counts
        for (x, y, w, _) in boxes:
            crop = np.ascontiguousarray(frame[y:y + w, x:x + w, :])
            if nets[0] == "reference":
                nets[2](crop)
            else:
                O.hand_call(nets[2], crop, backend="lib")
    return time.perf_counter() - t0, nets


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on all host threads, one frame of the workload
    per step: the unmodified reference when its files are at hand (cpu_reference_frame), else the oracle port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the reference picks its device by torch.cuda.is_available() (body.py:31,59): hide the GPUs so that this arm is its
    # CPU path on the box's host cores (must happen before torch is imported, which is why torch imports live in functions)
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
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
            "config": {"workload": workload_name(wl, WORKLOADS[wl][4]),
                       "timing": "host wall clock around each step (CPU only, no device work); a step here is ONE frame of the "
                                 "workload (the GPU arm's step is the whole batch): frames/s is per-frame work either way"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": nets[0], "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


class ChunkList(object):
    """A step's frames held as device chunks [k,H,W,3]; slicing at chunk boundaries hands the chunks back."""

    def __init__(self, chunks):
        self.chunks = chunks
        self.n = sum(int(c.shape[0]) for c in chunks)

    def __len__(self):
        return self.n

    def __getitem__(self, sl):
        pos = 0
        for c in self.chunks:
            if pos == sl.start:
                return c
            pos += int(c.shape[0])
        raise IndexError(sl)


def measure(ctx, wl, B, steps, warmup, chunk=None, headline=False):
    """All legs of one workload on this rank; returns the dict rank 0 turns into a JSON line / sub_result."""
    import torch
    import torch.distributed as dist

    import isl_b200
    from isl_b200 import _lib, features, synth
    from isl_b200.extract import KeypointExtractor, shard_indices

    rank, world, local = ctx["rank"], ctx["world"], ctx["local"]
    mt, H, W, boxes, _ = WORKLOADS[wl]
    clips = wl == "C4"
    body = isl_b200.Body(synth.make_flat_weights(mt, seed=0, init=WEIGHT_INIT), mt, scale_search=SCALES)
    hand = isl_b200.Hand(synth.make_flat_weights("hand", seed=0, init=WEIGHT_INIT))
    ex = KeypointExtractor(body, hand, chunk=chunk)
    hand_boxes = [boxes] * B
    flush = ctx["flush"]
    L = _lib.lib()

    def frames_for(step):
        if clips:   # clip c, frame t -> seed 1000*c + t; this rank's frames of `world` clips
            return [synth.synth_frame(H, W, 1000 * ((step * world * B + i) // B) + i % B) for i in shard_indices(world * B, rank, world)]
        return [synth.synth_frame(H, W, (rank * 100003 + step * B + i) % (2 ** 31)) for i in range(B)]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(batches_of, n_steps):
        """K steps through the two-lane pipeline; returns (device ms between the brackets, results of the last step)."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        last = None
        if clips:
            # the reference's extraction loop (extract_features.py:143-173): this rank's frames of the K steps' clips in
            # batches through ONE pipeline (as a video is processed), one feature row per frame, then the host gather to
            # rank 0 and the merge in frame order - all inside the timed region
            def all_batches():
                ck_ = chunk or B
                for s in range(n_steps):
                    fr = batches_of(s)
                    for a in range(0, len(fr), ck_):
                        flush.zero_()
                        yield fr[a:a + ck_], [boxes] * len(fr[a:a + ck_])
            barrier()
            e0.record()
            own = [s * world * B + i for s in range(n_steps) for i in shard_indices(world * B, rank, world)]
            rows, results = [], []
            for rs in ex.pipeline(all_batches()):   # the rows of a batch are built while the GPU works on the next one
                for (c, sb, hp) in rs:
                    # every worker keeps its own rows, as the reference's workers do (extract_features_mp.py:142-147
                    # saveFeaturesDict per process); what travels to rank 0 is the extractor's result per frame
                    rows.append(features.feature_record(c, sb, hp, frame_no=own[len(rows)], model_type=mt))
                    results.append((c, sb, hp))
            if world > 1:
                gathered = [None] * world if rank == 0 else None
                dist.gather_object((own, results), gathered, dst=0)
            else:
                gathered = [(own, results)]
            if rank == 0:
                last = [None] * (n_steps * world * B)
                for idxs, rws in gathered:
                    for i, r in zip(idxs, rws):
                        last[i] = r
                assert all(r is not None for r in last)
            e1.record()
            barrier()
            return e0.elapsed_time(e1), last

        def feed():
            for s in range(n_steps):
                flush.zero_()   # evicts L2 between steps (the per-step working set is far larger than L2 anyway)
                yield batches_of(s), hand_boxes
        barrier()
        e0.record()
        for res in ex.pipeline(feed()):
            last = res
        e1.record()
        barrier()
        return e0.elapsed_time(e1), last

    out = {"workload": workload_name(wl, B), "frames_per_step_per_rank": B}
    # ---- leg 1: device-resident inputs (value) ---------------------------------------------------------------
    n_distinct = min(warmup + steps, 4 if B * H * W < 40e6 else 2)
    host_sets = [frames_for(s) for s in range(n_distinct)]
    if clips:
        dev_sets = [[torch.from_numpy(f).cuda() for f in hs] for hs in host_sets]   # run_sharded indexes single frames
        dev_sets = [[torch.stack(ds[a:a + (chunk or B)]) for a in range(0, len(ds), chunk or B)] for ds in dev_sets]
    else:
        dev_sets = [torch.from_numpy(np.stack(hs)).cuda() for hs in host_sets]

    if clips:
        # device-resident variant: the same sharded batches, already in HBM
        def run_steps_dev(n_steps, first):
            return run_steps(lambda s: ChunkList(dev_sets[(first + s) % len(dev_sets)]), n_steps)
        run_steps_dev(warmup, 0)
    else:
        run_steps(lambda s: dev_sets[s % len(dev_sets)], warmup)   # W untimed steps (both lanes get their buffers)
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = L.islpose_launch_count()
    if clips:
        dev_ms, _ = run_steps_dev(steps, warmup)
    else:
        dev_ms, _ = run_steps(lambda s: dev_sets[(warmup + s) % len(dev_sets)], steps)
    out["gpu_launches"] = int(L.islpose_launch_count() - launches0)
    sampler.stop_flag = True
    out["clocks"] = sampler.summary()

    # ---- leg 2: host API end to end (e2e) ------------------------------------------------------------------
    # a step's frames wait in pinned host memory, as a frame feeder leaves them (frames.FrameFeeder decodes straight into
    # pinned batch buffers); the timed region holds their H2D copy, all device work and the D2H of every result
    pinned_sets = [torch.from_numpy(np.stack(hs)).pin_memory() for hs in host_sets]
    run_steps(lambda s: pinned_sets[s % len(pinned_sets)], 2)
    e2e_ms, res = run_steps(lambda s: pinned_sets[s % len(pinned_sets)], steps)
    if clips:
        out["d2h"] = int(sum(c.nbytes + sb.nbytes + sum(p.nbytes for p in hp) for c, sb, hp in res)) // max(steps, 1) if res else 0
    else:
        out["d2h"] = int(sum(c.nbytes + sb.nbytes + sum(p.nbytes for p in hp) for c, sb, hp in res))
    out["h2d"] = B * H * W * 3   # hand crops are cut from the device copy of the frame

    # ---- leg 3: the convolution plans of one step on their own (roofline of the tcgen05 kernels) -------------
    # exactly the network replays a step performs (same chunks, lanes and streams; resize and first-layer launches
    # included), without the post-processing, so the events time tensor-bound work only
    ck = ex.chunk or B
    lanes = ex._lane_streams(torch)
    main = torch.cuda.current_stream()
    flat_dev = torch.cat(dev_sets[0]) if clips else dev_sets[0]

    def networks_only(frames_dev):
        flops = 0
        ready = torch.cuda.Event()
        ready.record(main)
        for ci, a in enumerate(range(0, B, ck)):
            sub = frames_dev[a:a + ck]
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

    networks_only(flat_dev)
    barrier()
    conv_events, conv_flops = [], 0
    conv_sampler = ClockSampler(local)
    conv_sampler.start()
    for s in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        conv_flops += networks_only(flat_dev)
        e1.record()
        conv_events.append((e0, e1))
    barrier()
    conv_sampler.stop_flag = True
    conv_ms = sum(a.elapsed_time(b) for a, b in conv_events)
    out["conv_clocks"] = conv_sampler.summary()

    # ---- leg 4: the body map post-processing on its own (HBM roofline of subsystem 2) -------------------------
    # maps accumulation + gaussian/NMS/peak lists for one chunk, from the network outputs left by leg 3
    nb = min(ck, B)
    ws = body._workspace(nb, H, W, 0)
    maps = body.network_outputs(flat_dev[:nb], H, W, lane=0)
    parts = body.njoint - 1
    heat_scales = body._scales_struct(maps, 1)
    need = L.islpose_maps_workspace_floats(heat_scales, len(maps), nb, parts)
    if ws.get("mid") is None or ws["mid"].numel() < need:
        ws["mid"] = torch.empty((need,), dtype=torch.float32, device="cuda")

    def post_maps():
        _lib.check(L.islpose_maps_accumulate(heat_scales, len(maps), body.njoint, nb, H, W, parts, 1, _lib.ptr(ws["heat"]),
                                             _lib.ptr(ws["mid"]), ws["mid"].numel(), _lib.stream_ptr()), "maps_accumulate")
        _lib.check(L.islpose_body_peaks(_lib.ptr(ws["heat"]), nb * parts, H, W, body._gauss, body.thre1, ws["cap"],
                                        _lib.ptr(ws["counts"]), _lib.ptr(ws["keys"]), _lib.ptr(ws["scores"]),
                                        _lib.ptr(ws["overflow"]), _lib.stream_ptr()), "body_peaks")

    post_maps()
    barrier()
    post_events = []
    for s in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        post_maps()
        e1.record()
        post_events.append((e0, e1))
    barrier()
    post_ms = sum(a.elapsed_time(b) for a, b in post_events) / steps
    ws["overflow"].zero_()
    # SURVEY.md section 8d convention for the bytes of the reference dataflow, heat maps only (the PAF maps are never
    # materialised here; their share of B_post is reported separately as what the lazy sampler avoids)
    S = len(SCALES)
    grid_bytes = sum(4 * (m[2][2] // 8) * (m[2][3] // 8) for m in maps)
    out["post_bytes_heat"] = nb * (4 * H * W * ((2 * S - 1) * body.njoint + 2 * parts) + grid_bytes * body.njoint)
    out["post_bytes_all"] = nb * (4 * H * W * ((2 * S - 1) * (body.njoint + body.npaf) + 2 * parts) + grid_bytes * (body.njoint + body.npaf))
    out["post_frames"] = nb

    # ---- single-frame latency through the same public API (BASELINE.json's C2 is literally one frame) -------
    single_ms = None
    if headline:
        one = [host_sets[0][0]]
        for _ in range(2):
            ex.batch(one, hand_boxes[:1])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            ex.batch(one, hand_boxes[:1])
        torch.cuda.synchronize()
        single_ms = (time.perf_counter() - t0) / 5 * 1e3
    out["single_ms"] = single_ms

    times = torch.tensor([dev_ms, e2e_ms, conv_ms, post_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    out["dev_ms"], out["e2e_ms"], out["conv_ms"], out["post_ms"] = [float(x) for x in times.cpu()]
    out["conv_flops"] = conv_flops
    out["steps"], out["warmup"] = steps, warmup
    del body, hand, ex, dev_sets, flat_dev, ws, maps
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return out


def summarise(m, world, n_sm):
    """Measurement dict of one workload -> the keys of a JSON line (also the shape of a sub_result)."""
    sustained, burst, hbm_peak, src = peaks()
    B, steps = m["frames_per_step_per_rank"], m["steps"]
    frames_total = world * B * steps
    achieved = m["conv_flops"] / (m["conv_ms"] * 1e-3) / 1e12 if m["conv_ms"] > 0 else 0.0
    traffic, traffic_src = conv_dram_traffic()
    mhz = (m["conv_clocks"] or {}).get("sm_mhz")
    # the chip's own ceiling at the clock the leg actually ran at: 8192 dense bf16 FLOP per clock per SM
    # (build/mma_rate: one M=128 x N=256 x K=16 tcgen05.mma per 128 cycles, profiles/r1_mma_rate.txt)
    at_clock = n_sm * 8192 * mhz * 1e6 / 1e12 if mhz else None
    nb = m["post_frames"]
    post_gbs = m["post_bytes_heat"] / (m["post_ms"] * 1e-3) / 1e9
    return {
        "value": frames_total / (m["dev_ms"] * 1e-3), "unit": "frames/s", "ms_per_step": m["dev_ms"] / steps, "steps": steps,
        "warmup": m["warmup"],
        "e2e": {"value": frames_total / (m["e2e_ms"] * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": m["h2d"],
                "d2h_bytes_per_step": m["d2h"]},
        "gpu_launches": m["gpu_launches"], "clocks": m["clocks"],
        "roofline": {"bound": "tensor", "kernel": "conv_umma_* (tcgen05 implicit-GEMM variants): the network replays of a step run "
                                                  "on their own, same chunks / lanes / streams as the step; the resize and first-"
                                                  "layer launches are inside the same events",
                     "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                     "peak_source": "%s bf16_tflops_sustained (burst %.1f)" % (src, burst),
                     "frac_at_clock": achieved / at_clock if at_clock else None,
                     "peak_at_clock": at_clock, "sm_mhz_during_leg": mhz,
                     "traffic": traffic, "traffic_source": traffic_src,
                     "conv_share_of_step": m["conv_ms"] / max(m["dev_ms"], 1e-9)},
        "post_roofline": {"bound": "hbm", "kernel": "upsample8 + resize_accumulate + gauss_nms + sort_peaks "
                                                    "(body maps of one %d-frame chunk, run on their own)" % nb,
                          "achieved": post_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": post_gbs / hbm_peak,
                          "limiter": "instruction issue and the FP64 pipe, not HBM: the reference's float32 (no FMA, fixed order) "
                                     "cubic stages and float64 accumulation / gaussian are reproduced bit for bit (DESIGN.md section 7)",
                          "ms_per_frame": m["post_ms"] / nb, "algorithmic_bytes_per_frame": m["post_bytes_heat"] // nb,
                          "reference_dataflow_bytes_per_frame_incl_paf": m["post_bytes_all"] // nb},
    }


def measure_classifier(torch, n_windows=1024, reps=20):
    """The sign classifier behind the key points (demo_isl_translate.py:72-99): windows of 20 x 156 feature rows
    per second through isl_b200.Translator, one launch per batch of windows, device-timed; and one window end to end."""
    from isl_b200 import translate as TR
    tr = TR.Translator(TR.random_weights(167, seed=0))
    rng = np.random.RandomState(0)
    win = torch.from_numpy(rng.uniform(0, 720, (n_windows, 20, 156))).cuda()
    for _ in range(3):
        tr(win)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        p = tr(win)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    one = win[:1].cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(20):
        tr.top(one)
    single_ms = (time.perf_counter() - t0) / 20 * 1e3
    return {"windows_per_s": n_windows / ms * 1e3, "ms_per_launch": ms, "windows_per_launch": n_windows, "classes": 167,
            "single_window_host_to_host_ms": single_ms, "weights": "seeded random (the trained model does not ship with the reference)",
            "probability_sum_check": float(p[0].sum())}


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
    ap.add_argument("--no-sub", action="store_true", help="headline workload only (no sub_results)")
    ap.add_argument("--sub-steps", type=int, default=4, help="timed steps of every sub_result workload")
    ap.add_argument("--weights", default="torch", choices=["torch", "he"], help="random-init distribution (he = noisy-map stress)")
    args = ap.parse_args()
    global WEIGHT_INIT
    WEIGHT_INIT = args.weights
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    # exactly ONE line goes to stdout: libraries that print there (NCCL's version banner under torchrun) are sent to
    # stderr while the measurement runs; the JSON line is written to the real stdout at the end
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist

    import isl_b200
    isl_b200.configure()   # 32 hardware queues for the pipeline's streams; must precede the CUDA context

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = {"rank": rank, "world": world, "local": local,
           "flush": torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")}  # > 126 MB of L2
    n_sm = torch.cuda.get_device_properties(local).multi_processor_count

    def chunk_of(wl, B):
        return args.chunk or (15 if wl == "C4" else None)

    wl = args.workload
    B = args.batch or WORKLOADS[wl][4]
    head = measure(ctx, wl, B, args.steps, args.warmup, chunk=chunk_of(wl, B), headline=True)
    subs = []
    if not args.no_sub:
        for other in ("C2", "C3", "C4", "C5"):
            if other == wl:
                continue
            Bo = WORKLOADS[other][4]
            subs.append((other, measure(ctx, other, Bo, args.sub_steps, 3, chunk=chunk_of(other, Bo))))

    classifier = measure_classifier(torch) if rank == 0 else None

    if rank == 0:
        mt, H, W, boxes, _ = WORKLOADS[wl]
        line = {"metric": METRIC, "n_gpus": world, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic"}
        line.update(summarise(head, world, n_sm))
        line["config"] = {"workload": head["workload"],
                          "weights": "seeded random init, %s (no trained weights ship with the reference)" % (
                              "nn.Conv2d default distribution" if WEIGHT_INIT == "torch" else "He-uniform (noisy maps: thousands of peaks)"),
                          "l2": "flushed with a 256 MiB write before every timed step",
                          "timing": "CUDA events on the launching stream around the K steps; max over ranks",
                          "pipeline": "steps run through KeypointExtractor.pipeline(): two lanes, the post-processing and copies of "
                                      "one step overlap the convolutions of the next; all K steps start and end inside the timed region",
                          "single_frame_latency_ms": round(head["single_ms"], 2) if head["single_ms"] else None}
        line["classifier"] = classifier
        line["sub_results"] = []
        for name, m in subs:
            sr = {"config": {"workload": m["workload"]}, "n_gpus": world}
            sr.update(summarise(m, world, n_sm))
            line["sub_results"].append(sr)
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample of the same workload on the host cores: the reference arm of this script in a child process
            # without GPUs (one untimed frame for weights and thread pools, then two timed frames: ~10-30 s of CPU work)
            child = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", wl,
                                    "--steps", "2", "--warmup", "1", "--weights", WEIGHT_INIT],
                                   capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
            rows = [l for l in child.stdout.splitlines() if l.startswith("{")]
            if child.returncode == 0 and rows:
                line["cpu_baseline"] = json.loads(rows[-1])["cpu_baseline"]
            else:
                line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": os.cpu_count() or 1, "kind": "port",
                                        "sample": "reference arm failed: " + child.stderr[-300:]}
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
