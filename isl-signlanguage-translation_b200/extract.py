"""Frame-level extraction: body -> handDetect -> hand -> offset peaks, per frame or per batch, sharded by rank.

The per-frame sequence is the reference's canonical one (demo.py:21-43, ISL_Model_parameter.py:51-60):

    candidate, subset = body_estimation(oriImg)
    for x, y, w, is_left in util.handDetect(candidate, subset, oriImg):
        peaks = hand_estimation(oriImg[y:y+w, x:x+w, :])
        peaks[:, 0] = np.where(peaks[:, 0] == 0, peaks[:, 0], peaks[:, 0] + x)     # 0 means "not found"
        peaks[:, 1] = np.where(peaks[:, 1] == 0, peaks[:, 1], peaks[:, 1] + y)

Frames are independent, so a video shards by frame index across ranks (one process per GPU) and results are
gathered on the host in frame order; there is no collective on the data path.
"""
import numpy as np

from . import util


def shard_indices(n_items, rank, world_size):
    """Frame indices owned by `rank`: i with i % world_size == rank (the order inside a shard is ascending)."""
    return list(range(rank, n_items, world_size))


def merge_shards(shards, n_items):
    """Inverse of shard_indices: shards[r] holds the results of rank r in its own order."""
    world = len(shards)
    out = [None] * n_items
    for r, part in enumerate(shards):
        idx = shard_indices(n_items, r, world)
        if len(idx) != len(part):
            raise ValueError("rank %d returned %d results for %d frames" % (r, len(part), len(idx)))
        for i, v in zip(idx, part):
            out[i] = v
    return out


class KeypointExtractor(object):
    """(candidate, subset, all_hand_peaks) per frame, like ISLSignPos.call (ISL_Model_parameter.py:51-60)."""

    def __init__(self, body, hand=None, chunk=4):
        self.body = body
        self.hand = hand
        self.chunk = chunk      # frames per pipeline stage of batch_device(); None = no pipelining
        self._lanes = None

    def __call__(self, frame):
        return self.batch([frame])[0]

    def batch(self, frames, hand_boxes=None):
        """frames: same-size uint8 [H,W,3] arrays. hand_boxes: optional per-frame list of [x, y, w, is_left] that
        replaces util.handDetect (benchmark configs fix the boxes because random-init weights find no persons).
        The frames are uploaded once; hand crops are cut from the device copy."""
        if hasattr(self.body, "upload"):
            return self.batch_device(self.body.upload(frames), hand_boxes)
        bodies = self.body.batch(frames)   # stand-in estimators without a device path (tests)
        if self.hand is None:
            return [(c, s, []) for c, s in bodies]
        crops, owner = [], []
        for fi, ((cand, sub), frame) in enumerate(zip(bodies, frames)):
            boxes = hand_boxes[fi] if hand_boxes is not None else util.handDetect(cand, sub, frame)
            for (x, y, w, is_left) in boxes:
                crops.append(np.asarray(frame)[y:y + w, x:x + w, :])
                owner.append((fi, x, y))
        peaks = self.hand.batch(crops) if crops else []
        return self._assemble(bodies, owner, peaks)

    @staticmethod
    def _assemble(bodies, owner, peaks):
        per_frame = [[] for _ in bodies]
        for (fi, x, y), p in zip(owner, peaks):
            p = p.copy()
            p[:, 0] = np.where(p[:, 0] == 0, p[:, 0], p[:, 0] + x)   # demo.py:36-37: 0 means "not found"
            p[:, 1] = np.where(p[:, 1] == 0, p[:, 1], p[:, 1] + y)
            per_frame[fi].append(p)
        return [(c, s, per_frame[i]) for i, (c, s) in enumerate(bodies)]

    def batch_device(self, frames_dev, hand_boxes, chunk=None):
        """Same as batch() for frames already resident on the device (uint8 cuda tensor [n,H,W,3]). hand_boxes=None
        runs util.handDetect on every frame's (candidate, subset) as the reference loop does. With `chunk` (or
        self.chunk) set, the batch runs through pipeline() in chunks of that many frames."""
        n = int(frames_dev.shape[0])
        chunk = int(chunk or self.chunk or n)
        if not hasattr(self.body, "enqueue") or n <= chunk:
            return self._batch_device_serial(frames_dev, hand_boxes)
        parts = [(frames_dev[a:a + chunk], None if hand_boxes is None else hand_boxes[a:a + chunk]) for a in range(0, n, chunk)]
        out = []
        for res in self.pipeline(parts):
            out.extend(res)
        return out

    def pipeline(self, batches, with_features=False):
        """Software pipeline over a sequence of batches: yields the result list of every batch, in order.
        with_features: every frame's result is (candidate, subset, hand_peaks, row) with `row` the classifier's float64
        [156] feature vector, formed on the device (csrc/features.cu) instead of (candidate, subset, hand_peaks).

        batches: iterable of (frames, hand_boxes); frames is a uint8 cuda tensor [n,H,W,3], a list of numpy frames
        (uploaded through pinned memory) or a pinned host tensor [n,H,W,3] (frames.FrameFeeder batches: copied as they
        are); hand_boxes a per-frame list of boxes or None (= util.handDetect).

        Two lanes (independent buffers and streams) alternate: while the host waits for the body results of batch i
        (it needs them for util.handDetect), the body networks of batch i+1 are already queued, and the hand networks
        of batch i are queued before the host waits for batch i+1. The memory- and latency-bound stages (map
        accumulation, peaks, grouping, hand key points, the copies in both directions) therefore run under the
        tensor-bound convolutions of the neighbouring batches. The per-frame dependency body -> handDetect -> hand
        is unchanged, and so are the results."""
        import torch

        lanes = self._lane_streams(torch)
        main = torch.cuda.current_stream()
        it = iter(batches)
        crops_cut = [None, None]   # per lane: the hand crops of the lane's previous batch have been copied out of its staging buffer

        def start_body(idx):
            try:
                frames, boxes = next(it)
            except StopIteration:
                return None
            if not (torch.is_tensor(frames) and frames.is_cuda):
                # the lane's staging buffer is about to be overwritten: its last readers are the body kernels of batch
                # idx - 2 (finished: their results were collected) and the crop copies of that batch on the hand stream
                frames = self.body.upload(frames, lane=idx % 2, after=crops_cut[idx % 2])
            ready = torch.cuda.Event()
            ready.record(main)
            st = lanes[idx % 2]
            with torch.cuda.stream(st):
                st.wait_event(ready)
                ticket = self.body.enqueue(frames, lane=idx % 2)
            return frames, boxes, ticket, ready

        def finish_hand(pending):
            bodies, owner, ticket = pending
            if self.hand is None:
                res = [(c, s, []) for c, s in bodies]
            else:
                res = self._assemble(bodies, owner, self.hand.finish(ticket))
            if with_features:
                ticket["done"].synchronize()
                rows = ticket["features"].numpy()
                res = [(c, s, hp, rows[i].copy()) for i, (c, s, hp) in enumerate(res)]
            return res

        idx = 0
        pending = None
        nxt = start_body(0)
        while nxt is not None:
            frames, boxes, ticket, ready = nxt
            nxt = start_body(idx + 1)
            bodies = self.body.finish(ticket)
            owner, hticket = [], None
            if self.hand is None and with_features:
                with torch.cuda.stream(lanes[2 + idx % 2]):
                    hticket = self._rows_to_host(torch, ticket["features"])
            if self.hand is not None:
                crops = []
                st = lanes[2 + idx % 2]
                with torch.cuda.stream(st):
                    st.wait_event(ready)
                    for fi, (cand, sub) in enumerate(bodies):
                        fb = boxes[fi] if boxes is not None else util.handDetect(cand, sub, frames[fi])
                        for (x, y, w, is_left) in fb:
                            crops.append(frames[fi, y:y + w, x:x + w, :].contiguous())
                            owner.append((fi, x, y))
                    cut = torch.cuda.Event()
                    cut.record(st)
                    crops_cut[idx % 2] = cut
                    feats = (ticket["features"], owner) if with_features else None
                    hticket = self.hand.enqueue(crops, lane=idx % 2, features=feats) if (crops or feats) else None
            if pending is not None:
                yield finish_hand(pending)
            pending = (bodies, owner, hticket)
            idx += 1
        if pending is not None:
            yield finish_hand(pending)
        for st in lanes:
            main.wait_stream(st)

    @staticmethod
    def _rows_to_host(torch, rows):
        host = torch.empty(rows.shape, dtype=rows.dtype).pin_memory()
        host.copy_(rows, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())
        return dict(done=done, features=host, keep=rows)

    def _lane_streams(self, torch):
        if self._lanes is None:
            # body lanes run the post-processing on their own (high-priority) stream: its short kernels are on the
            # critical path of the host loop and must not queue behind thousands of convolution CTAs
            self._lanes = [torch.cuda.Stream(device=self.body.device, priority=-1) for _ in range(2)] + \
                          [torch.cuda.Stream(device=self.body.device) for _ in range(2)]
        return self._lanes

    def _batch_device_serial(self, frames_dev, hand_boxes):
        bodies = self.body.batch_device(frames_dev)
        if self.hand is None:
            return [(c, s, []) for c, s in bodies]
        crops, owner = [], []
        for fi, (cand, sub) in enumerate(bodies):
            boxes = hand_boxes[fi] if hand_boxes is not None else util.handDetect(cand, sub, frames_dev[fi])
            for (x, y, w, is_left) in boxes:
                crops.append(frames_dev[fi, y:y + w, x:x + w, :].contiguous())
                owner.append((fi, x, y))
        peaks = self.hand.batch_device(crops) if crops else []
        return self._assemble(bodies, owner, peaks)

    def features(self, frames, batch_size=8):
        """Clip -> float64 [T,156]: the per-frame feature vectors of demo_isl_translate.py's loop, frames processed in
        batches through pipeline(); the rows are formed on the device and arrive with the key points."""
        frames = list(frames)
        batches = [(frames[a:a + batch_size], None) for a in range(0, len(frames), batch_size)]
        if not hasattr(self.body, "enqueue"):   # stand-in estimators without a device path (tests)
            from . import features as F
            mt = getattr(self.body, "model_type", "coco")
            rows = [F.frame_features(c, s, hp, mt) for b, _ in batches for (c, s, hp) in self.batch(b)]
        else:
            rows = [r for res in self.pipeline(batches, with_features=True) for (_, _, _, r) in res]
        return np.stack(rows) if rows else np.zeros((0, 156))

    def translate(self, frames, translator, batch_size=8):
        """Clip -> (class index, probability) per frame from the 21st on: the whole of demo_isl_translate.py's loop
        (:176-197), with every frame's key points extracted once instead of once per window it appears in."""
        p = translator.sliding(self.features(frames, batch_size)).cpu().numpy()
        idx = p.argmax(axis=1) if len(p) else np.zeros((0,), np.int64)
        return [(int(i), float(p[k, i])) for k, i in enumerate(idx)]

    def records(self, frames, batch_size=8, first_frame_no=0, **meta):
        """Clip -> list of per-frame feature rows (features.feature_record = extract_features.py's saveFeature dict)."""
        from . import features as F

        frames = list(frames)
        mt = getattr(self.body, "model_type", "coco")
        batches = [(frames[a:a + batch_size], None) for a in range(0, len(frames), batch_size)]
        runner = self.pipeline(batches) if hasattr(self.body, "enqueue") else (self.batch(b) for b, _ in batches)
        rows = []
        for res in runner:
            for (c, s, hp) in res:
                rows.append(F.feature_record(c, s, hp, frame_no=first_frame_no + len(rows), model_type=mt, **meta))
        return rows

    def run_sharded(self, frames, rank, world_size, batch_size=8, hand_boxes=None):
        """Processes this rank's shard of `frames` (any sequence indexable by global frame index) in batches through
        pipeline(); returns results in shard order (merge_shards() of all ranks' lists restores frame order)."""
        idx = shard_indices(len(frames), rank, world_size)
        sels = [idx[b:b + batch_size] for b in range(0, len(idx), batch_size)]
        gen = (([frames[i] for i in sel], None if hand_boxes is None else [hand_boxes[i] for i in sel]) for sel in sels)
        out = []
        if hasattr(self.body, "enqueue"):
            for res in self.pipeline(gen):
                out.extend(res)
        else:
            for fr, hb in gen:
                out.extend(self.batch(fr, hb))
        return out
