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

    def __init__(self, body, hand=None):
        self.body = body
        self.hand = hand

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

    def batch_device(self, frames_dev, hand_boxes):
        """Same as batch() for frames already resident on the device (uint8 cuda tensor [n,H,W,3]); the hand boxes
        must be given because util.handDetect needs the host copy of candidate/subset either way."""
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

    def run_sharded(self, frames, rank, world_size, batch_size=8, hand_boxes=None):
        """Processes this rank's shard of `frames` in batches; returns results in shard order."""
        idx = shard_indices(len(frames), rank, world_size)
        out = []
        for b in range(0, len(idx), batch_size):
            sel = idx[b:b + batch_size]
            out.extend(self.batch([frames[i] for i in sel], None if hand_boxes is None else [hand_boxes[i] for i in sel]))
        return out
