"""Hand estimator with the reference's call API (src/hand.py), running on the sm_100a kernels.

    hand = Hand(model_path)                  # hand.py:16
    peaks = hand(crop)                       # hand.py:24   crop: uint8 [h,w,3] -> int array [21,2] (x, y), [0,0] = missing
    all_peaks = hand.batch([crop0, ...])     # new: the four network scales run once for all crops
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .body import _load_flat, scale_geometry
from .nets import PoseNet
from .util import gaussian_weights


def _pow2_at_least(n):
    p = 1
    while p < n:
        p <<= 1
    return p


class Hand(object):
    def __init__(self, model_path, device=None, tuning=None):
        self.model = PoseNet('hand', _load_flat(model_path), device=device, tuning=tuning)
        self.device = self.model.device
        self.scale_search = [0.5, 1.0, 1.5, 2.0]   # hand.py:25
        self.boxsize, self.stride, self.padValue, self.thre = 368, 8, 128, 0.05
        self._gauss = (C.c_double * 25)(*gaussian_weights().tolist())
        self._streams = {}   # lane -> side streams

    def __call__(self, oriImg):
        return self.batch([oriImg])[0]

    def network_outputs(self, crops_dev, lane=0):
        """crops_dev: list of uint8 cuda tensors [h,w,3]. Crops whose network inputs have the same shape (all square
        crops do: 184, 368, 552, 736) share a batched plan per scale. Returns per crop a list of
        (heat tensor, plane offset in elements, geometry)."""
        L = _lib.lib()
        geoms = [scale_geometry(c.shape[0], c.shape[1], self.scale_search, self.boxsize) for c in crops_dev]
        per_crop = [[] for _ in crops_dev]
        # plan instances are created (and their buffers zero-filled) before the streams fork
        group_sizes = {}
        for g in geoms:
            for si in range(len(self.scale_search)):
                key = (si, g[si][3], g[si][4])
                group_sizes[key] = group_sizes.get(key, 0) + 1
        for (si, hp, wp), cnt in group_sizes.items():
            self.model.instance(_pow2_at_least(cnt), hp, wp, lane)
        streams = self._streams.setdefault(lane, [])
        main = torch.cuda.current_stream()
        timing = self.model.timing
        if timing is not None:
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record(main)
        fork = torch.cuda.Event()
        fork.record(main)
        flops = launches = 0
        used = 0
        for si in range(len(self.scale_search)):
            groups = {}
            for ci, g in enumerate(geoms):
                groups.setdefault((g[si][3], g[si][4]), []).append(ci)
            for (hp, wp), members in groups.items():
                # every (scale, input shape) group is an independent network replay: one stream each
                while len(streams) <= used:
                    streams.append(torch.cuda.Stream(device=self.device))
                side = streams[used]
                used += 1
                inst = self.model.instance(_pow2_at_least(len(members)), hp, wp, lane)
                with torch.cuda.stream(side):
                    side.wait_event(fork)
                    for slot, ci in enumerate(members):
                        m, rh, rw, _, _ = geoms[ci][si]
                        c = crops_dev[ci]
                        dst = C.c_void_p(inst.input.data_ptr() + slot * 3 * hp * wp * 4)
                        _lib.check(L.islpose_resize_pad_normalize(_lib.ptr(c), 1, c.shape[0], c.shape[1], m, rh, rw, hp, wp,
                                                                  dst, None, _lib.stream_ptr()), "islpose_resize_pad_normalize")
                    inst.run()
                    done = torch.cuda.Event()
                    done.record(side)
                    main.wait_event(done)
                flops += inst.flops_algorithmic * len(members) // inst.n
                launches += inst.launches + len(members)
                plane = 22 * (hp // 8) * (wp // 8)
                for slot, ci in enumerate(members):
                    per_crop[ci].append((inst.outputs[0], slot * plane, (geoms[ci][si][1], geoms[ci][si][2], hp, wp)))
        if timing is not None:
            t1.record(main)
            timing.append((t0, t1, flops, launches))
        return per_crop

    def postprocess(self, maps, h, w):
        """maps: [(heat tensor, element offset, (rh, rw, hp, wp))] for one crop -> int array [21,2]."""
        L = _lib.lib()
        st = _lib.stream_ptr()
        dev = self.device
        arr = (_lib.Scale * len(maps))()
        for i, (t, off, (rh, rw, hp, wp)) in enumerate(maps):
            arr[i].lowres = t.data_ptr() + 4 * off
            arr[i].gh, arr[i].gw, arr[i].hc, arr[i].wc = hp // 8, wp // 8, rh, rw
        heat = torch.empty((21, h, w), dtype=torch.float64, device=dev)
        smoothed = torch.empty((21, h, w), dtype=torch.float64, device=dev)
        labels = torch.empty((21, h, w), dtype=torch.int32, device=dev)
        mass = torch.empty((21, h, w), dtype=torch.float64, device=dev)
        out = torch.zeros((21, 2), dtype=torch.int32, device=dev)
        # small crops: one pass (fewer launches); large crops: materialise the up-sampled maps first (less arithmetic)
        mid = None
        if h * w >= 256 * 256:
            mid = torch.empty((L.islpose_maps_workspace_floats(arr, len(maps), 1, 21),), dtype=torch.float32, device=dev)
        _lib.check(L.islpose_maps_accumulate(arr, len(maps), 22, 1, h, w, 21, 0, _lib.ptr(heat), _lib.ptr(mid),
                                             mid.numel() if mid is not None else 0, st), "islpose_maps_accumulate")
        _lib.check(L.islpose_hand_peaks(_lib.ptr(heat), 21, h, w, self._gauss, self.thre, _lib.ptr(smoothed),
                                        _lib.ptr(labels), _lib.ptr(mass), _lib.ptr(out), st), "islpose_hand_peaks")
        return out

    def batch(self, crops):
        crops = [np.ascontiguousarray(c) for c in crops]
        for c in crops:
            if c.ndim != 3 or c.shape[2] != 3 or c.dtype != np.uint8:
                raise ValueError("Hand needs uint8 [h,w,3] crops, got %s %s" % (c.shape, c.dtype))
        if not crops:
            return []
        with torch.cuda.device(self.device):
            return self.batch_device([torch.from_numpy(c).to(self.device, non_blocking=True) for c in crops])

    MAX_CROPS_PER_REPLAY = 32   # crops per network replay: bounds the plan buffers (2.2 GB per 64-channel full-resolution
                                # buffer at 32 x 736 x 736); more crops run as consecutive replays of the same plans

    def enqueue(self, dev_crops, lane=0):
        """Launches the four network scales and the key-point selection of every crop (list of contiguous uint8
        cuda tensors [h,w,3]) without waiting; finish(ticket) returns the list of int64 [21,2] arrays."""
        if not dev_crops:
            return None
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            outs = []
            for a in range(0, len(dev_crops), self.MAX_CROPS_PER_REPLAY):
                part = dev_crops[a:a + self.MAX_CROPS_PER_REPLAY]
                per_crop = self.network_outputs(part, lane)
                # a crop's post-processing launches only 21 CTAs per kernel: spread the crops over a few streams; the
                # join below also keeps the next replay from overwriting network outputs that are still being read
                fork = torch.cuda.Event()
                fork.record(main)
                lanes = min(len(part), 8)
                streams = self._streams.setdefault(lane, [])
                while len(streams) < lanes:
                    streams.append(torch.cuda.Stream(device=self.device))
                for i in range(len(part)):
                    side = streams[i % lanes] if lanes > 1 else main
                    with torch.cuda.stream(side):
                        side.wait_event(fork)
                        outs.append(self.postprocess(per_crop[i], part[i].shape[0], part[i].shape[1]))
                if lanes > 1:
                    for j in range(lanes):
                        done = torch.cuda.Event()
                        done.record(streams[j])
                        main.wait_event(done)
            stacked = torch.stack(outs)
            host = torch.empty(stacked.shape, dtype=stacked.dtype).pin_memory()
            host.copy_(stacked, non_blocking=True)
            done = torch.cuda.Event()
            done.record(main)
        return dict(host=host, done=done, n=len(dev_crops), keep=(outs, stacked, dev_crops))

    def finish(self, ticket):
        if ticket is None:
            return []
        ticket["done"].synchronize()
        stacked = ticket["host"].numpy()
        return [stacked[i].astype(np.int64) for i in range(ticket["n"])]

    def batch_device(self, dev_crops):
        """dev_crops: list of contiguous uint8 cuda tensors [h,w,3] -> list of int64 arrays [21,2]."""
        return self.finish(self.enqueue(dev_crops))
