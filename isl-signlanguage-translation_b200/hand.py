"""Hand estimator with the reference's call API (src/hand.py), running on the sm_100a kernels.

    hand = Hand(model_path)                  # hand.py:16
    peaks = hand(crop)                       # hand.py:24   crop: uint8 [h,w,3] -> int array [21,2] (x, y), [0,0] = missing
    all_peaks = hand.batch([crop0, ...])     # new: the four network scales run once for all crops
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .body import scale_geometry
from .nets import PoseNet
from .util import gaussian_weights
from .weights import load_flat


def _pow2_at_least(n):
    p = 1
    while p < n:
        p <<= 1
    return p


class Hand(object):
    MAX_CROPS_PER_REPLAY = 32   # crops per network replay: bounds the plan buffers (2.2 GB per 64-channel full-resolution
                                # buffer at 32 x 736 x 736); more crops run as consecutive replays of the same plans

    def __init__(self, model_path, device=None, tuning=None):
        self.model = PoseNet('hand', load_flat(model_path), device=device, tuning=tuning)
        self.device = self.model.device
        self.scale_search = [0.5, 1.0, 1.5, 2.0]   # hand.py:25
        self.boxsize, self.stride, self.padValue, self.thre = 368, 8, 128, 0.05
        self._gauss = (C.c_double * 25)(*gaussian_weights().tolist())
        self._streams = {}   # lane -> side streams
        self._work = {}      # lane -> (uint8 scratch tensor for the key-point kernels, int32 results)

    def __call__(self, oriImg):
        return self.batch([oriImg])[0]

    def network_outputs(self, crops_dev, lane=0):
        """crops_dev: list of uint8 cuda tensors [h,w,3]. Crops whose network inputs have the same shape (all square
        crops do: 184, 368, 552, 736) share one batched replay per scale, sized for exactly that many crops: the plan
        records its launches over the leading images of a power-of-two sized set of buffers, so no padded image is
        ever computed. Returns per crop a list of (heat tensor, plane offset in elements, geometry)."""
        L = _lib.lib()
        geoms = [scale_geometry(c.shape[0], c.shape[1], self.scale_search, self.boxsize) for c in crops_dev]
        per_crop = [[] for _ in crops_dev]
        # plan instances are created (and their buffers zero-filled) before the streams fork
        groups_of = []
        for si in range(len(self.scale_search)):
            groups = {}
            for ci, g in enumerate(geoms):
                groups.setdefault((g[si][3], g[si][4]), []).append(ci)
            groups_of.append(groups)
        # all (scale, shape) groups run side by side on their own streams: a few crops do not fill the device, so every
        # group keeps to a share of the SMs (PoseNet.share_sms)
        flat_groups = [(si, hw, members) for si in range(len(self.scale_search)) for hw, members in groups_of[si].items()]
        shares = self.model.share_sms([(len(m), hw[0], hw[1]) for (_, hw, m) in flat_groups])
        budget = {(si, hw): b for (si, hw, _), b in zip(flat_groups, shares)}
        for (si, hw, members) in flat_groups:
            self.model.instance(len(members), hw[0], hw[1], lane, exact_of=_pow2_at_least(len(members)), sm_budget=budget[(si, hw)])
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
            for (hp, wp), members in groups_of[si].items():
                # every (scale, input shape) group is an independent network replay: one stream each
                while len(streams) <= used:
                    streams.append(torch.cuda.Stream(device=self.device))
                side = streams[used]
                used += 1
                inst = self.model.instance(len(members), hp, wp, lane, exact_of=_pow2_at_least(len(members)),
                                           sm_budget=budget[(si, (hp, wp))])
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
                flops += inst.flops_algorithmic
                launches += inst.launches + len(members)
                plane = 22 * (hp // 8) * (wp // 8)
                for slot, ci in enumerate(members):
                    per_crop[ci].append((inst.outputs[0], slot * plane, (geoms[ci][si][1], geoms[ci][si][2], hp, wp)))
        if timing is not None:
            t1.record(main)
            timing.append((t0, t1, flops, launches))
        return per_crop

    def keypoints(self, per_crop, sizes, out, lane=0):
        """per_crop[i]: [(heat tensor, element offset, (rh, rw, hp, wp))] per scale, sizes[i] = (h, w) of crop i;
        out: int32 cuda tensor [len(sizes), 21, 2] that receives (x, y). Three launches for up to 32 crops, on the
        current stream, inside one scratch buffer per lane."""
        L = _lib.lib()
        n = len(sizes)
        crops = (_lib.HandCrop * n)()
        for i, (maps, (h, w)) in enumerate(zip(per_crop, sizes)):
            crops[i].h, crops[i].w = h, w
            for s, (t, off, (rh, rw, hp, wp)) in enumerate(maps):
                sc = crops[i].scales[s]
                sc.lowres = t.data_ptr() + 4 * off
                sc.gh, sc.gw, sc.hc, sc.wc = hp // 8, wp // 8, rh, rw
        need = L.islpose_hand_workspace_bytes(crops, n)
        key = (lane, torch.cuda.current_stream().cuda_stream)   # launches on one stream are ordered: one scratch per stream
        ws = self._work.get(key)
        if ws is None or ws.numel() < need:
            # the previous (smaller) buffer may still be in use by launches in flight on this lane: let the caching
            # allocator keep it alive for them (record_stream) instead of synchronising
            if ws is not None:
                ws.record_stream(torch.cuda.current_stream())
            ws = self._work[key] = torch.empty((int(need * 1.25) + 256,), dtype=torch.uint8, device=self.device)
        _lib.check(L.islpose_hand_keypoints(crops, n, len(self.scale_search), self._gauss, self.thre, _lib.ptr(ws), ws.numel(),
                                            _lib.ptr(out), _lib.stream_ptr()), "islpose_hand_keypoints")

    def postprocess(self, maps, h, w):
        """maps: [(heat tensor, element offset, (rh, rw, hp, wp))] for one crop -> int32 cuda tensor [21,2]."""
        out = torch.zeros((1, 21, 2), dtype=torch.int32, device=self.device)
        self.keypoints([maps], [(h, w)], out)
        return out[0]

    def batch(self, crops):
        crops = [np.ascontiguousarray(c) for c in crops]
        for c in crops:
            if c.ndim != 3 or c.shape[2] != 3 or c.dtype != np.uint8:
                raise ValueError("Hand needs uint8 [h,w,3] crops, got %s %s" % (c.shape, c.dtype))
        if not crops:
            return []
        with torch.cuda.device(self.device):
            return self.batch_device([torch.from_numpy(c).to(self.device, non_blocking=True) for c in crops])

    def enqueue(self, dev_crops, lane=0, features=None):
        """Launches the four network scales and the key-point selection of every crop (list of contiguous uint8
        cuda tensors [h,w,3]) without waiting; finish(ticket) returns the list of int64 [21,2] arrays.
        features = (rows, owner): `rows` is the float64 cuda tensor [frames,156] whose body part Body wrote, owner[i] =
        (frame, crop x, crop y) of crop i in util.handDetect order; the hand parts of the rows are filled on the device
        and the finished rows travel to the host with the key points (ticket["features"])."""
        if not dev_crops:
            if features is None:
                return None
            with torch.cuda.device(self.device):
                return self._features_only(features[0])
        with torch.cuda.device(self.device):
            main = torch.cuda.current_stream()
            out = torch.zeros((len(dev_crops), 21, 2), dtype=torch.int32, device=self.device)
            for a in range(0, len(dev_crops), self.MAX_CROPS_PER_REPLAY):
                part = dev_crops[a:a + self.MAX_CROPS_PER_REPLAY]
                per_crop = self.network_outputs(part, lane)
                # the key-point kernels run on the main stream, so the next replay (whose network streams fork from
                # main) cannot overwrite network outputs that are still being read
                self.keypoints(per_crop, [(c.shape[0], c.shape[1]) for c in part], out[a:a + len(part)], lane)
            host = torch.empty(out.shape, dtype=out.dtype).pin_memory()
            host.copy_(out, non_blocking=True)
            feat_host = None
            if features is not None:
                rows, owner = features
                seen, table = {}, []
                for (fi, x, y) in owner:
                    slot = seen.get(fi, 0)
                    seen[fi] = slot + 1
                    table.append((fi, slot if slot < 2 else -1, x, y))   # util.get_handpose keeps two hands (util.py:198)
                tab = torch.tensor(table, dtype=torch.int32).pin_memory().to(self.device, non_blocking=True)
                _lib.check(_lib.lib().islpose_hand_features(_lib.ptr(tab), _lib.ptr(out), len(table), rows.shape[0],
                                                            _lib.ptr(rows), _lib.stream_ptr()), "islpose_hand_features")
                feat_host = torch.empty(rows.shape, dtype=rows.dtype).pin_memory()
                feat_host.copy_(rows, non_blocking=True)
                out = (out, tab)
            done = torch.cuda.Event()
            done.record(main)
        return dict(host=host, done=done, n=len(dev_crops), keep=(out, dev_crops), features=feat_host)

    def _features_only(self, rows):
        """A batch without hand crops: the rows (body part only) still travel to the host."""
        feat_host = torch.empty(rows.shape, dtype=rows.dtype).pin_memory()
        feat_host.copy_(rows, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())
        return dict(host=None, done=done, n=0, keep=(rows,), features=feat_host)

    def finish(self, ticket):
        if ticket is None:
            return []
        ticket["done"].synchronize()
        if ticket["host"] is None:
            return []
        stacked = ticket["host"].numpy()
        return [stacked[i].astype(np.int64) for i in range(ticket["n"])]

    def batch_device(self, dev_crops):
        """dev_crops: list of contiguous uint8 cuda tensors [h,w,3] -> list of int64 arrays [21,2]."""
        return self.finish(self.enqueue(dev_crops))
