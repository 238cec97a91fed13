"""Times every phase of a step on its own (CUDA events, L2 flushed before each repetition), so that the phases can be
compared with their rooflines without a profiler:

    python tools/phase_times.py [C2|C3] [frames]

Phases: body networks (each scale alone, then all scales concurrently), map accumulation, gaussian + NMS + peak
lists, PAF scoring + matching + assembly, hand networks, hand key points. Set ISLPOSE_GAUSS=1 to time the
first-generation gaussian kernel instead of the sliding-window one.
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import isl_b200  # noqa: E402
isl_b200.configure()
from isl_b200 import _lib, synth  # noqa: E402
from isl_b200.body import scale_geometry  # noqa: E402


def timed(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    reps = 5
    mt, H, W, boxes, _ = bench.WORKLOADS[wl]
    torch.cuda.set_device(0)
    L = _lib.lib()
    body = isl_b200.Body(synth.make_flat_weights(mt, seed=0, init="torch"), mt, scale_search=bench.SCALES)
    hand = isl_b200.Hand(synth.make_flat_weights("hand", seed=0, init="torch"))
    frames = torch.from_numpy(np.stack([synth.synth_frame(H, W, i) for i in range(nb)])).cuda()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    sustained, burst, hbm, _ = bench.peaks()
    print("== %s, %d frames" % (wl, nb))

    # ---- body networks
    geoms = scale_geometry(H, W, bench.SCALES, 368)
    total_flops = 0
    for (m, rh, rw, hp, wp) in geoms:
        inst = body.model.instance(nb, hp, wp, 0)

        def one(inst=inst, m=m, rh=rh, rw=rw, hp=hp, wp=wp):
            _lib.check(L.islpose_resize_pad_normalize(_lib.ptr(frames), nb, H, W, m, rh, rw, hp, wp, _lib.ptr(inst.input), None,
                                                      _lib.stream_ptr()), "resize")
            inst.run()
        ms = timed(one, reps, flush)
        total_flops += inst.flops_algorithmic
        print("body net %4dx%-4d alone      %8.3f ms  %7.1f TFLOP/s  (%d launches)" % (hp, wp, ms, inst.flops_algorithmic / ms / 1e9,
                                                                                  inst.launches))
    ms = timed(lambda: body.network_outputs(frames, H, W, 0), reps, flush)
    print("body nets, 4 scales together %8.3f ms  %7.1f TFLOP/s = %.1f%% of %.0f" % (ms, total_flops / ms / 1e9,
                                                                               100 * total_flops / ms / 1e9 / sustained, sustained))

    # ---- body post-processing
    ws = body._workspace(nb, H, W, 0)
    maps = body.network_outputs(frames, H, W, 0)
    parts = body.njoint - 1
    heat_scales = body._scales_struct(maps, 1)
    need = L.islpose_maps_workspace_floats(heat_scales, len(maps), nb, parts)
    ws["mid"] = torch.empty((need,), dtype=torch.float32, device="cuda")
    S = len(maps)
    grid_bytes = sum(4 * (m[2][2] // 8) * (m[2][3] // 8) for m in maps)
    acc_bytes = nb * (4 * H * W * (2 * S - 1) * body.njoint + grid_bytes * body.njoint)
    peak_bytes = nb * 4 * H * W * 2 * parts

    def accumulate():
        _lib.check(L.islpose_maps_accumulate(heat_scales, S, body.njoint, nb, H, W, parts, 1, _lib.ptr(ws["heat"]),
                                             _lib.ptr(ws["mid"]), ws["mid"].numel(), _lib.stream_ptr()), "acc")

    def accumulate_single():
        _lib.check(L.islpose_maps_accumulate(heat_scales, S, body.njoint, nb, H, W, parts, 1, _lib.ptr(ws["heat"]),
                                             None, 0, _lib.stream_ptr()), "acc1")

    def peaks():
        _lib.check(L.islpose_body_peaks(_lib.ptr(ws["heat"]), nb * parts, H, W, body._gauss, body.thre1, ws["cap"],
                                        _lib.ptr(ws["counts"]), _lib.ptr(ws["keys"]), _lib.ptr(ws["scores"]),
                                        _lib.ptr(ws["overflow"]), _lib.stream_ptr()), "peaks")

    ms = timed(accumulate, reps, flush)
    print("maps accumulate (two pass)   %8.3f ms  %7.1f GB/s (SURVEY 8d bytes) = %.1f%% of %.0f" % (
        ms, acc_bytes / ms / 1e6, 100 * acc_bytes / ms / 1e6 / hbm, hbm))
    ms1 = timed(accumulate_single, 2, flush)
    print("maps accumulate (one pass)   %8.3f ms" % ms1)
    accumulate()
    ms = timed(peaks, reps, flush)
    print("gaussian + NMS + sort        %8.3f ms  %7.1f GB/s (SURVEY 8d bytes) = %.1f%% of %.0f" % (
        ms, peak_bytes / ms / 1e6, 100 * peak_bytes / ms / 1e6 / hbm, hbm))
    counts = ws["counts"].cpu().numpy().reshape(nb, parts)
    print("   peaks per frame: %s" % counts.sum(axis=1).tolist())

    body.post_finish(body.post_enqueue(maps, nb, H, W, ws))   # allocates the result staging buffers

    def group():
        body._group(maps, nb, H, W, ws)
    try:
        ms = timed(group, reps, flush)
        torch.cuda.synchronize()
        tail = ws["tail_host"].numpy() if ws.get("tail_host") is not None else None
        print("PAF score + match + assemble %8.3f ms   overflow %s" % (ms, None if tail is None else int(tail[-1])))
    except Exception as e:  # noqa: BLE001
        print("group failed:", e)
    ws["overflow"].zero_()

    # ---- hand
    crops = [frames[i, y:y + w, x:x + w, :].contiguous() for i in range(nb) for (x, y, w, _) in boxes]
    hand.model.timing = []
    ms = timed(lambda: hand.network_outputs(crops, 0), reps, flush)
    flops = hand.model.timing[-1][2]
    hand.model.timing = None
    print("hand nets, %2d crops x 4 scales %7.3f ms  %7.1f TFLOP/s = %.1f%% of %.0f" % (len(crops), ms, flops / ms / 1e9,
                                                                                 100 * flops / ms / 1e9 / sustained, sustained))
    per_crop = hand.network_outputs(crops, 0)

    kp = torch.zeros((len(crops), 21, 2), dtype=torch.int32, device="cuda")

    def hand_post():
        for a in range(0, len(crops), 32):
            hand.keypoints(per_crop[a:a + 32], [(c.shape[0], c.shape[1]) for c in crops[a:a + 32]], kp[a:a + 32])
    ms = timed(hand_post, reps, flush)
    print("hand key points, %2d crops, batched %7.3f ms" % (len(crops), ms))
    ms = timed(lambda: hand.finish(hand.enqueue(crops, 0)), reps, flush)
    print("hand enqueue+finish (nets + post)     %7.3f ms" % ms)


if __name__ == "__main__":
    main()
