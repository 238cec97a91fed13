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

PEAK_CAP = 1024     # peaks per (frame, part)
CAND_CAP = 2048     # accepted connection candidates per (frame, limb)
MAX_PERSON = 512


def _load_flat(model_path):
    if isinstance(model_path, dict):
        return model_path
    return torch.load(model_path, map_location="cpu")


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
        self.model = PoseNet(kind, _load_flat(model_path), device=device, tuning=tuning)
        self.device = self.model.device
        self._gauss = (C.c_double * 25)(*gaussian_weights().tolist())
        self._work = {}
        self.last_overflow = 0

    # ------------------------------------------------------------------------------------------------
    def __call__(self, oriImg):
        return self.batch([oriImg])[0]

    def _workspace(self, n, H, W):
        key = (n, H, W)
        ws = self._work.get(key)
        if ws is None:
            dev = self.device
            parts = self.njoint - 1
            nl = 24 if self._kind == 'body25' else 19
            i32 = dict(dtype=torch.int32, device=dev)
            f64 = dict(dtype=torch.float64, device=dev)
            ws = dict(
                frames=torch.empty((n, H, W, 3), dtype=torch.uint8, device=dev),
                heat=torch.empty((n, parts, H, W), **f64),
                counts=torch.zeros((n * parts,), **i32),
                keys=torch.zeros((n * parts, PEAK_CAP), dtype=torch.int32, device=dev),
                scores=torch.zeros((n * parts, PEAK_CAP), **f64),
                cand_count=torch.zeros((n * nl,), **i32),
                cand_pair=torch.zeros((n * nl, CAND_CAP), **i32),
                cand_score=torch.zeros((n * nl, CAND_CAP), **f64),
                conn_count=torch.zeros((n * nl,), **i32),
                conn_ij=torch.zeros((n * nl, PEAK_CAP, 2), **i32),
                conn_score=torch.zeros((n * nl, PEAK_CAP), **f64),
                candidate=torch.zeros((n, parts * PEAK_CAP, 4), **f64),
                n_cand=torch.zeros((n,), **i32),
                subset=torch.zeros((n, MAX_PERSON, self.njoint + 1), **f64),
                n_person=torch.zeros((n,), **i32),
                overflow=torch.zeros((1,), **i32),
                pinned=torch.empty((n, H, W, 3), dtype=torch.uint8).pin_memory(),
            )
            self._work[key] = ws
        return ws

    def network_outputs(self, frames_dev, H, W):
        """Runs every scale; returns [(paf, heat, geometry)] with the plans' float32 NCHW output tensors."""
        L = _lib.lib()
        n = frames_dev.shape[0]
        outs = []
        for (m, rh, rw, hp, wp) in scale_geometry(H, W, self.scale_search, self.boxsize):
            inst = self.model.instance(n, hp, wp)
            _lib.check(L.islpose_resize_pad_normalize(_lib.ptr(frames_dev), n, H, W, m, rh, rw, hp, wp,
                                                      _lib.ptr(inst.input), None, _lib.stream_ptr()),
                       "islpose_resize_pad_normalize")
            inst.run()
            outs.append((inst.outputs[0], inst.outputs[1], (rh, rw, hp, wp)))
        return outs

    def _scales_struct(self, maps, which):
        arr = (_lib.Scale * len(maps))()
        for i, m in enumerate(maps):
            t = m[which]
            rh, rw, hp, wp = m[2]
            arr[i].lowres = t.data_ptr()
            arr[i].gh, arr[i].gw, arr[i].hc, arr[i].wc = hp // 8, wp // 8, rh, rw
        return arr

    def postprocess(self, maps, n, H, W, ws):
        """Peaks, PAF scoring and grouping from the per-scale network outputs -> list of (candidate, subset)."""
        L = _lib.lib()
        st = _lib.stream_ptr()
        parts = self.njoint - 1
        heat_scales = self._scales_struct(maps, 1)
        paf_scales = self._scales_struct(maps, 0)
        _lib.check(L.islpose_maps_accumulate(heat_scales, len(maps), self.njoint, n, H, W, parts, 1, _lib.ptr(ws["heat"]), st),
                   "islpose_maps_accumulate")
        _lib.check(L.islpose_body_peaks(_lib.ptr(ws["heat"]), n * parts, H, W, self._gauss, self.thre1, PEAK_CAP,
                                        _lib.ptr(ws["counts"]), _lib.ptr(ws["keys"]), _lib.ptr(ws["scores"]),
                                        _lib.ptr(ws["overflow"]), st), "islpose_body_peaks")
        gb = _lib.GroupBuffers()
        gb.cap, gb.cand_cap, gb.max_cand, gb.max_person = PEAK_CAP, CAND_CAP, parts * PEAK_CAP, MAX_PERSON
        for f in ("counts", "keys", "scores", "cand_count", "cand_pair", "cand_score", "conn_count", "conn_ij",
                  "conn_score", "candidate", "n_cand", "subset", "n_person", "overflow"):
            setattr(gb, f, ws[f].data_ptr())
        _lib.check(L.islpose_body_group(paf_scales, len(maps), 1 if self._kind == 'body25' else 0, n, H, W, self.thre2,
                                        self.mid_num, C.byref(gb), st), "islpose_body_group")
        n_cand = ws["n_cand"].cpu().numpy()
        n_person = ws["n_person"].cpu().numpy()
        self.last_overflow = int(ws["overflow"].cpu().item())
        if self.last_overflow:
            ws["overflow"].zero_()
            raise _lib.IslposeError("body grouping exceeded a fixed capacity (code %d: peaks per part > %d, candidates per "
                                    "limb > %d or persons > %d)" % (self.last_overflow, PEAK_CAP, CAND_CAP, MAX_PERSON))
        max_c, max_p = int(n_cand.max()) if n else 0, int(n_person.max()) if n else 0
        cand = ws["candidate"][:, :max(max_c, 1)].cpu().numpy()
        sub = ws["subset"][:, :max(max_p, 1)].cpu().numpy()
        results = []
        for i in range(n):
            c = cand[i, :n_cand[i]].copy() if n_cand[i] else np.array([])   # body.py:183 gives shape (0,) when empty
            s = sub[i, :n_person[i]].copy() if n_person[i] else -1 * np.ones((0, self.njoint + 1))
            results.append((c, s))
        return results

    def batch(self, frames):
        """frames: list of uint8 [H,W,3] BGR arrays of one size -> list of (candidate, subset)."""
        frames = [np.asarray(f) for f in frames]
        H, W = frames[0].shape[:2]
        for f in frames:
            if f.shape != (H, W, 3) or f.dtype != np.uint8:
                raise ValueError("Body.batch needs uint8 [H,W,3] frames of one size, got %s %s" % (f.shape, f.dtype))
        n = len(frames)
        with torch.cuda.device(self.device):
            ws = self._workspace(n, H, W)
            host = ws["pinned"].numpy()
            for i, f in enumerate(frames):
                host[i] = f   # also resolves negative-stride views such as frame[:, :, ::-1]
            ws["frames"].copy_(ws["pinned"], non_blocking=True)
            maps = self.network_outputs(ws["frames"], H, W)
            return self.postprocess(maps, n, H, W, ws)
