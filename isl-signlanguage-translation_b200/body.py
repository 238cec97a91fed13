"""Body estimator with the reference's call API (src/body.py), running on the sm_100a kernels.

    body = Body(model_path, model_type='coco')        # body.py:16
    candidate, subset = body(oriImg)                   # body.py:39   oriImg: uint8 [H,W,3] BGR
    results = body.batch([frame0, frame1, ...])        # new: same-size frames through one batched launch chain

Defaults are the reference's (scale_search=[0.5], boxsize 368, thre1 0.1, thre2 0.05, body.py:40-46); the
4-scale list the reference keeps commented out (body.py:40) is selected with scale_search=[0.5, 1.0, 1.5, 2.0].
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .nets import PoseNet
from .tables import model_dims
from .util import gaussian_weights
from .weights import load_flat

PEAK_CAP = 1024        # initial peaks per (frame, part); grows (x2) up to MAX_PEAK_CAP when a frame needs more
MAX_PEAK_CAP = 4096    # the sort / matching kernels' limit (csrc/prepost.cuh kMaxPeakCap)
PAIR_CAP = 128 * 1024  # initial nA*nB capacity per (frame, limb); grows (x4) when a frame needs more
MAX_PERSON = 4096      # initial row slots per frame (rows ever created); grows (x4) up to 65536


def scale_geometry(H, W, scale_search, boxsize):
    """Per scale: (multiplier, resized rh x rw, padded hp x wp) as body.py:47,53-54 derive them."""
    out = []
    for s in scale_search:
        m = s * boxsize / H
        rh, rw = int(np.rint(H * m)), int(np.rint(W * m))
        out.append((m, rh, rw, (rh + 7) // 8 * 8, (rw + 7) // 8 * 8))
    return out


class Body(object):
    def __init__(self, model_path, model_type='coco', scale_search=None, boxsize=368, thre1=0.1, thre2=0.05,
                 device=None, tuning=None):
        if model_type not in ('coco', 'body25'):
            print('not right model_type, use coco')   # body.py:25-29 falls back the same way
            kind = 'coco'
        else:
            kind = model_type
        self.njoint, self.npaf = model_dims(kind)
        self.model_type = model_type
        self._kind = kind
        self.scale_search = list(scale_search) if scale_search is not None else [0.5]
        self.boxsize, self.stride, self.padValue = boxsize, 8, 128
        self.thre1, self.thre2, self.mid_num = thre1, thre2, 10
        self.model = PoseNet(kind, load_flat(model_path), device=device, tuning=tuning)
        self.device = self.model.device
        self._gauss = (C.c_double * 25)(*gaussian_weights().tolist())
        self._work = {}
        self._staging = {}   # (n, H, W) -> (pinned host frames, device frames)
        self._streams = {}   # lane -> side streams, one per scale
        self.last_overflow = 0

    # ------------------------------------------------------------------------------------------------
    def __call__(self, oriImg):
        return self.batch([oriImg])[0]

    def _workspace(self, n, H, W, lane=0):
        key = (n, H, W, lane)
        ws = self._work.get(key)
        if ws is None:
            dev = self.device
            parts = self.njoint - 1
            nl = 24 if self._kind == 'body25' else 19
            i32 = dict(dtype=torch.int32, device=dev)
            f64 = dict(dtype=torch.float64, device=dev)
            ws = dict(
                heat=torch.empty((n, parts, H, W), **f64),
                counts=torch.zeros((n * parts,), **i32),
                pair_cap=PAIR_CAP,
                pair_score=torch.empty((n * nl, PAIR_CAP), **f64),
                conn_count=torch.zeros((n * nl,), **i32),
                max_person=MAX_PERSON,
                n_cand=torch.zeros((n,), **i32),
                subset=torch.zeros((n, MAX_PERSON, self.njoint + 1), **f64),
                n_person=torch.zeros((n,), **i32),
                overflow=torch.zeros((1,), **i32),
            )
            self._size_peak_buffers(ws, n, PEAK_CAP)
            self._work[key] = ws
        return ws

    def _size_peak_buffers(self, ws, n, cap):
        """(Re)allocates everything whose size follows the peak capacity per (frame, part)."""
        dev = self.device
        parts = self.njoint - 1
        nl = 24 if self._kind == 'body25' else 19
        i32 = dict(dtype=torch.int32, device=dev)
        f64 = dict(dtype=torch.float64, device=dev)
        ws.update(
            cap=cap,
            keys=torch.zeros((n * parts, cap), **i32),
            scores=torch.zeros((n * parts, cap), **f64),
            end_paf=torch.empty((n * nl, 2, cap, 2), **f64),
            conn_ij=torch.zeros((n * nl, cap, 2), **i32),
            conn_score=torch.zeros((n * nl, cap), **f64),
            owner=torch.zeros((n, parts * cap, 2), **i32),
            candidate=torch.zeros((n, parts * cap, 4), **f64),
        )

    def network_outputs(self, frames_dev, H, W, lane=0):
        """Runs every scale; returns [(paf, heat, geometry)] with the plans' float32 NCHW output tensors.
        The scales are independent (body.py:51 loops over them sequentially), so each runs on its own stream: the
        small-scale networks are latency-bound (a few dozen CTAs per layer) and hide under the large ones.
        `lane` selects an independent set of plan buffers and streams, so that two batches can be in flight."""
        L = _lib.lib()
        n = frames_dev.shape[0]
        geoms = scale_geometry(H, W, self.scale_search, self.boxsize)
        budgets = self.model.share_sms([(n, hp, wp) for (_, _, _, hp, wp) in geoms])
        insts = [self.model.instance(n, hp, wp, lane, sm_budget=b) for (_, _, _, hp, wp), b in zip(geoms, budgets)]
        main = torch.cuda.current_stream()
        streams = self._streams.setdefault(lane, [])
        while len(streams) < len(geoms):
            streams.append(torch.cuda.Stream(device=self.device))
        timing = self.model.timing
        if timing is not None:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(main)
        fork = torch.cuda.Event()
        fork.record(main)
        outs = []
        for i, ((m, rh, rw, hp, wp), inst) in enumerate(zip(geoms, insts)):
            side = streams[i] if len(geoms) > 1 else main
            with torch.cuda.stream(side):
                side.wait_event(fork)
                _lib.check(L.islpose_resize_pad_normalize(_lib.ptr(frames_dev), n, H, W, m, rh, rw, hp, wp,
                                                          _lib.ptr(inst.input), None, _lib.stream_ptr()),
                           "islpose_resize_pad_normalize")
                inst.run()
                if side is not main:
                    done = torch.cuda.Event()
                    done.record(side)
                    main.wait_event(done)
            outs.append((inst.outputs[0], inst.outputs[1], (rh, rw, hp, wp)))
        if timing is not None:
            t1.record(main)
            timing.append((t0, t1, sum(i.flops_algorithmic for i in insts), sum(i.launches for i in insts) + len(insts)))
        return outs

    def _scales_struct(self, maps, which):
        arr = (_lib.Scale * len(maps))()
        for i, m in enumerate(maps):
            t = m[which]
            rh, rw, hp, wp = m[2]
            arr[i].lowres = t.data_ptr()
            arr[i].gh, arr[i].gw, arr[i].hc, arr[i].wc = hp // 8, wp // 8, rh, rw
        return arr

    def _peaks(self, n, H, W, ws):
        _lib.check(_lib.lib().islpose_body_peaks(_lib.ptr(ws["heat"]), n * (self.njoint - 1), H, W, self._gauss, self.thre1,
                                                 ws["cap"], _lib.ptr(ws["counts"]), _lib.ptr(ws["keys"]), _lib.ptr(ws["scores"]),
                                                 _lib.ptr(ws["overflow"]), _lib.stream_ptr()), "islpose_body_peaks")

    def _group(self, maps, n, H, W, ws):
        L = _lib.lib()
        parts = self.njoint - 1
        paf_scales = self._scales_struct(maps, 0)
        gb = _lib.GroupBuffers()
        gb.cap, gb.pair_cap, gb.max_cand, gb.max_person = ws["cap"], ws["pair_cap"], parts * ws["cap"], ws["max_person"]
        for f in ("counts", "keys", "scores", "pair_score", "end_paf", "conn_count", "conn_ij", "conn_score", "owner", "candidate",
                  "n_cand", "subset", "n_person", "overflow"):
            setattr(gb, f, ws[f].data_ptr())
        _lib.check(L.islpose_body_group(paf_scales, len(maps), 1 if self._kind == 'body25' else 0, n, H, W, self.thre2,
                                        self.mid_num, C.byref(gb), _lib.stream_ptr()), "islpose_body_group")
        # body half of the classifier's feature rows (util.get_bodypose circles), straight from the grouping outputs. A
        # fresh tensor per call: the hand stage of this batch fills its half while later batches run on this workspace
        ws["features"] = torch.empty((n, 156), dtype=torch.float64, device=self.device)
        _lib.check(L.islpose_body_features(_lib.ptr(ws["candidate"]), _lib.ptr(ws["subset"]), _lib.ptr(ws["n_person"]), n,
                                           parts * ws["cap"], ws["max_person"], 1 if self._kind == 'body25' else 0,
                                           _lib.ptr(ws["features"]), _lib.stream_ptr()), "islpose_body_features")
        # the three small result tables travel together, asynchronously, into pinned memory, and with them the leading
        # rows of candidate / subset (as many as recent calls needed, with head room): one synchronisation per call
        ws["tail_dev"][:n].copy_(ws["n_cand"])
        ws["tail_dev"][n:2 * n].copy_(ws["n_person"])
        ws["tail_dev"][2 * n:2 * n + 1].copy_(ws["overflow"])
        ws["tail_host"].copy_(ws["tail_dev"], non_blocking=True)
        self._copy_rows(ws, n)

    def _copy_rows(self, ws, n):
        rc, rp = ws.setdefault("rows_c", 512), ws.setdefault("rows_p", 64)
        rc, rp = min(rc, ws["candidate"].shape[1]), min(rp, ws["subset"].shape[1])
        if ws.get("cand_host") is None or ws["cand_host"].shape[1] != rc:
            ws["cand_host"] = torch.empty((n, rc, 4), dtype=torch.float64).pin_memory()
        if ws.get("sub_host") is None or ws["sub_host"].shape[1] != rp:
            ws["sub_host"] = torch.empty((n, rp, self.njoint + 1), dtype=torch.float64).pin_memory()
        ws["cand_host"].copy_(ws["candidate"][:, :rc], non_blocking=True)
        ws["sub_host"].copy_(ws["subset"][:, :rp], non_blocking=True)

    def post_enqueue(self, maps, n, H, W, ws):
        """Launches peaks, PAF scoring and grouping for the per-scale network outputs on the current stream (no
        host synchronisation); post_finish() collects the results."""
        L = _lib.lib()
        st = _lib.stream_ptr()
        parts = self.njoint - 1
        heat_scales = self._scales_struct(maps, 1)
        need = L.islpose_maps_workspace_floats(heat_scales, len(maps), n, parts)
        if ws.get("mid") is None or ws["mid"].numel() < need:
            ws["mid"] = torch.empty((need,), dtype=torch.float32, device=self.device)
        if ws.get("tail_dev") is None:
            ws["tail_dev"] = torch.zeros((2 * n + 1,), dtype=torch.int32, device=self.device)
            ws["tail_host"] = torch.zeros((2 * n + 1,), dtype=torch.int32).pin_memory()
        _lib.check(L.islpose_maps_accumulate(heat_scales, len(maps), self.njoint, n, H, W, parts, 1, _lib.ptr(ws["heat"]),
                                             _lib.ptr(ws["mid"]), ws["mid"].numel(), st), "islpose_maps_accumulate")
        self._peaks(n, H, W, ws)
        self._group(maps, n, H, W, ws)
        done = torch.cuda.Event()
        done.record()
        return dict(maps=maps, n=n, H=H, W=W, ws=ws, done=done, stream=torch.cuda.current_stream(), features=ws["features"])

    def post_finish(self, ticket):
        """Waits for a post_enqueue() ticket and returns the list of (candidate, subset)."""
        maps, n, H, W, ws = ticket["maps"], ticket["n"], ticket["H"], ticket["W"], ticket["ws"]
        with torch.cuda.stream(ticket["stream"]):
            while True:
                ticket["done"].synchronize()
                tail = ws["tail_host"].numpy()
                n_cand, n_person = tail[:n].copy(), tail[n:2 * n].copy()
                flags = self.last_overflow = int(tail[2 * n])
                max_c, max_p = int(n_cand.max()) if n else 0, int(n_person.max()) if n else 0
                redo = False
                if flags:
                    ws["overflow"].zero_()
                    # the peak lists were truncated (which peaks survive depends on the order of the atomics), so nothing
                    # computed from them is returned: larger lists, then peaks and grouping again from the heat maps,
                    # which are still in the workspace. The reference has no such limit; this one ends at MAX_PEAK_CAP.
                    if flags & (_lib.OVERFLOW_PEAKS | _lib.OVERFLOW_CANDIDATES):
                        if ws["cap"] >= MAX_PEAK_CAP:
                            raise _lib.IslposeError("a body part has more than %d peaks in one frame (overflow flags %d)" % (
                                MAX_PEAK_CAP, flags))
                        newcap = 1024
                        while newcap <= ws["cap"]:
                            newcap *= 2
                        self._size_peak_buffers(ws, n, min(newcap, MAX_PEAK_CAP))
                        self._peaks(n, H, W, ws)
                    if flags & _lib.OVERFLOW_PAIRS:
                        if ws["pair_cap"] >= ws["cap"] * ws["cap"]:
                            raise _lib.IslposeError("pair matrix overflow at its maximum size")
                        # a limb has more candidate pairs than the scratch matrix holds: enlarge it and redo the grouping
                        ws["pair_cap"] = min(ws["pair_cap"] * 4, ws["cap"] * ws["cap"])
                        ws["pair_score"] = torch.empty((ws["conn_count"].numel(), ws["pair_cap"]), dtype=torch.float64,
                                                       device=self.device)
                    if flags & _lib.OVERFLOW_PERSONS:
                        if ws["max_person"] >= 65536:
                            raise _lib.IslposeError("more than 65536 person rows in one frame")
                        ws["max_person"] *= 4
                        ws["subset"] = torch.zeros((n, ws["max_person"], self.njoint + 1), dtype=torch.float64,
                                                   device=self.device)
                    redo = True
                elif max_c > ws["cand_host"].shape[1] or max_p > ws["sub_host"].shape[1]:
                    # more rows than were copied speculatively: fetch them now and copy more next time
                    ws["rows_c"], ws["rows_p"] = max(ws["rows_c"], 2 * max_c), max(ws["rows_p"], 2 * max_p)
                    self._copy_rows(ws, n)
                    torch.cuda.current_stream().synchronize()
                if not redo:
                    break
                self._group(maps, n, H, W, ws)
                ticket["features"] = ws["features"]   # the rows of the repeated grouping, not of the truncated one
                ticket["done"] = torch.cuda.Event()
                ticket["done"].record()
            cand, sub = ws["cand_host"].numpy(), ws["sub_host"].numpy()
        results = []
        for i in range(n):
            c = cand[i, :n_cand[i]].copy() if n_cand[i] else np.array([])   # body.py:183 gives shape (0,) when empty
            s = sub[i, :n_person[i]].copy() if n_person[i] else -1 * np.ones((0, self.njoint + 1))
            results.append((c, s))
        return results

    def postprocess(self, maps, n, H, W, ws):
        """Peaks, PAF scoring and grouping from the per-scale network outputs -> list of (candidate, subset)."""
        return self.post_finish(self.post_enqueue(maps, n, H, W, ws))

    def upload(self, frames, lane=0, after=None):
        """Host frames -> this call's device staging buffer [n,H,W,3] (through pinned memory; one buffer per lane).
        `after`: an event the copy must wait for - the last reader of the lane's previous contents."""
        if torch.is_tensor(frames):
            # a (pinned) host tensor [n,H,W,3], e.g. a FrameFeeder batch: one asynchronous DMA, no host-side copy
            if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3 or not frames.is_contiguous():
                raise ValueError("Body.upload needs a contiguous uint8 [n,H,W,3] tensor, got %s %s" % (tuple(frames.shape), frames.dtype))
            key = (frames.shape[0], frames.shape[1], frames.shape[2], lane)
            stage = self._staging.get(key)
            if stage is None:
                stage = (None, torch.empty(tuple(frames.shape), dtype=torch.uint8, device=self.device))
                self._staging[key] = stage
            if after is not None:
                torch.cuda.current_stream().wait_event(after)
            stage[1].copy_(frames, non_blocking=True)
            return stage[1]
        frames = [np.asarray(f) for f in frames]
        H, W = frames[0].shape[:2]
        for f in frames:
            if f.shape != (H, W, 3) or f.dtype != np.uint8:
                raise ValueError("Body.batch needs uint8 [H,W,3] frames of one size, got %s %s" % (f.shape, f.dtype))
        key = (len(frames), H, W, lane)
        stage = self._staging.get(key)
        if stage is None or stage[0] is None:
            stage = (torch.empty((len(frames), H, W, 3), dtype=torch.uint8).pin_memory(),
                     stage[1] if stage is not None else torch.empty((len(frames), H, W, 3), dtype=torch.uint8, device=self.device))
            self._staging[key] = stage
        host = stage[0].numpy()
        for i, f in enumerate(frames):
            host[i] = f   # also resolves negative-stride views such as frame[:, :, ::-1]
        if after is not None:
            torch.cuda.current_stream().wait_event(after)
        stage[1].copy_(stage[0], non_blocking=True)
        return stage[1]

    def enqueue(self, frames_dev, lane=0):
        """Launches networks and post-processing for frames_dev (uint8 cuda tensor [n,H,W,3]) on the current stream
        and the lane's side streams without waiting for anything; finish(ticket) returns the results. Two lanes
        let the (memory / latency bound) post-processing of one batch run under the convolutions of the next."""
        n, H, W, _ = frames_dev.shape
        with torch.cuda.device(self.device):
            ws = self._workspace(n, H, W, lane)
            maps = self.network_outputs(frames_dev, H, W, lane)
            return self.post_enqueue(maps, n, H, W, ws)

    def finish(self, ticket):
        with torch.cuda.device(self.device):
            return self.post_finish(ticket)

    def batch_device(self, frames_dev):
        """frames_dev: uint8 cuda tensor [n,H,W,3] (contiguous) -> list of (candidate, subset)."""
        return self.finish(self.enqueue(frames_dev))

    def batch(self, frames):
        """frames: list of uint8 [H,W,3] BGR arrays of one size -> list of (candidate, subset)."""
        with torch.cuda.device(self.device):
            return self.batch_device(self.upload(frames))
