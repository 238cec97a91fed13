"""The frame loop either side of the key-point path: video file -> frames -> extractor ->
per-frame JSON + feature rows, as the reference's extract_features.py:143-173 / ISL_extract_features_videos.py:128-147 /
extract_features_mp.py:110-147 do, rebuilt around what a 200 frames/s/GPU extractor needs:

  FrameFeeder     a reader thread per process decodes this rank's share of a video (cv2.VideoCapture - the image has no
                  NVDEC SDK) straight into a ring of PINNED batch buffers, so the H2D copy of a batch is one asynchronous
                  DMA and decoding batch i+1 overlaps the GPU work of batch i
  VideoExtractor  is_processed / saveFeature / extract_features_worker with the reference's names and file layout;
                  frames whose JSON already exists are skipped without being decoded to pixels (resume by existence,
                  extract_features.py:97-101), JSON files are written by a writer thread

Frames shard across ranks in contiguous blocks (block_range): a decoder has to walk a file in order, so interleaved
sharding would make every rank decode every frame; results merge by frame index on the host, no collective.
"""
import json
import os
import queue
import threading
import time

import numpy as np

from . import features as F


def block_range(n_items, rank, world_size):
    """[start, stop) of rank's contiguous share of n_items (sizes differ by at most one)."""
    base, extra = divmod(n_items, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


class ArraySource(object):
    """Frames held in memory or in a np.memmap: uint8 [T,H,W,3] BGR (a raw video dump)."""

    def __init__(self, array):
        self.array = array
        self.pos = 0

    def __len__(self):
        return len(self.array)

    @property
    def shape(self):
        return tuple(self.array.shape[1:3])

    def seek(self, index):
        self.pos = index

    def skip(self):
        self.pos += 1

    def read_into(self, dst):
        dst[...] = self.array[self.pos]
        self.pos += 1
        return True


class VideoSource(object):
    """cv2.VideoCapture over a file; frames come out BGR like the reference hands them to Body (`frame[:, :, ::-1]` of
    an RGB reader, extract_features.py:163)."""

    def __init__(self, path):
        import cv2

        self.cv2 = cv2
        self.cap = cv2.VideoCapture(path)
        if not self.cap.isOpened():
            raise IOError("cannot open video %s" % path)
        self.n = int(self.cap.get(cv2.CAP_PROP_FRAME_COUNT))
        self.hw = (int(self.cap.get(cv2.CAP_PROP_FRAME_HEIGHT)), int(self.cap.get(cv2.CAP_PROP_FRAME_WIDTH)))

    def __len__(self):
        return self.n

    @property
    def shape(self):
        return self.hw

    def seek(self, index):
        if index:
            self.cap.set(self.cv2.CAP_PROP_POS_FRAMES, index)

    def skip(self):
        self.cap.grab()   # advances without colour conversion / copy

    def read_into(self, dst):
        ok, frame = self.cap.read()
        if not ok:
            return False
        dst[...] = frame
        return True


def open_source(src):
    if isinstance(src, (ArraySource, VideoSource)):
        return src
    if isinstance(src, np.ndarray):
        return ArraySource(src)
    path = os.fspath(src)
    if path.endswith(".npy"):
        return ArraySource(np.load(path, mmap_mode="r"))
    return VideoSource(path)


class FrameFeeder(object):
    """Iterates (frame indices, pinned uint8 tensor [n,H,W,3]) over this rank's block of a source's frames.

    A daemon thread fills `n_buffers` pinned batch buffers in turn; the consumer must be done with a batch's tensor (its
    H2D copy complete) before asking for the batch after the next one - KeypointExtractor.pipeline() guarantees that,
    it collects a batch's body results before it uploads the batch two positions later. `skip(index)` frames are
    advanced over without decoding to pixels."""

    def __init__(self, src, batch_size=8, rank=0, world_size=1, n_buffers=4, skip=None, limit=None):
        import torch

        self.src = open_source(src)
        total = len(self.src) if limit is None else min(len(self.src), limit)
        self.start, self.stop = block_range(total, rank, world_size)
        self.batch_size = batch_size
        H, W = self.src.shape
        self.buffers = [torch.empty((batch_size, H, W, 3), dtype=torch.uint8).pin_memory() if torch.cuda.is_available()
                        else torch.empty((batch_size, H, W, 3), dtype=torch.uint8) for _ in range(n_buffers)]
        self.free = queue.Queue()
        for i in range(n_buffers):
            self.free.put(i)
        self.ready = queue.Queue(maxsize=n_buffers)
        self.skip = skip or (lambda index: False)
        self.decode_s = 0.0
        self.frames_decoded = 0
        self.held = []
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        try:
            self.src.seek(self.start)
            slot, fill, idxs = None, 0, []
            for index in range(self.start, self.stop):
                if self.skip(index):
                    self.src.skip()
                    continue
                if slot is None:
                    slot, fill, idxs = self.free.get(), 0, []
                t0 = time.perf_counter()
                ok = self.src.read_into(self.buffers[slot][fill].numpy())
                self.decode_s += time.perf_counter() - t0
                if not ok:
                    break
                self.frames_decoded += 1
                idxs.append(index)
                fill += 1
                if fill == self.batch_size:
                    self.ready.put((slot, idxs))
                    slot = None
            if slot is not None and idxs:
                self.ready.put((slot, idxs))
            self.ready.put(None)
        except Exception as e:  # noqa: BLE001 - surfaced to the consumer
            self.ready.put(e)

    def __iter__(self):
        while True:
            item = self.ready.get()
            if item is None:
                return
            if isinstance(item, Exception):
                raise item
            slot, idxs = item
            # a buffer goes back to the reader two batches later (see the class docstring)
            self.held.append(slot)
            if len(self.held) > 2:
                self.free.put(self.held.pop(0))
            yield idxs, self.buffers[slot][:len(idxs)]


class VideoExtractor(object):
    """The reference's per-video worker (extract_features.py `ISLFeatureExtractor`): same method names, same files.

        vx = VideoExtractor(KeypointExtractor(body, hand), transforms_path_parent='out/transforms')
        rows = vx.extract_features_worker('videos/MVI_2978.MOV', 'Adjectives', 'loud')

    Per frame it writes `<parent>/<type>/<expression>/<stem>-original/<filename>-<idx>.json` holding candidate, subset
    and all_hand_peaks (extract_features.py:112-117) and returns the saveFeature rows. Divergence, documented: a frame
    counts as processed when its JSON exists; the reference also demands the preview JPG that only its test mode draws
    (extract_features.py:97-101), i.e. outside test mode it never resumes."""

    def __init__(self, extractor, transforms_path_parent, dataset_base_path="", model_type=None, batch_size=8, rank=0,
                 world_size=1, limit=None, hand_boxes=None):
        self.extractor = extractor
        self.transforms_path_parent = transforms_path_parent
        self.dataset_base_path = dataset_base_path
        self.model_type = model_type or getattr(extractor.body, "model_type", "body25")
        self.batch_size = batch_size
        self.rank, self.world_size = rank, world_size
        self.limit = limit
        self.hand_boxes = hand_boxes   # benchmarks only: fixed [x, y, w, is_left] boxes per frame instead of util.handDetect
        self.stats = {}

    def directory(self, filename, transform, label_type, label_expression):
        return os.path.join(self.transforms_path_parent, label_type, label_expression,
                            "%s-%s" % (filename.split('.')[0], transform))

    def is_processed(self, filename, idx, transform, label_type, label_expression):
        return os.path.exists(os.path.join(self.directory(filename, transform, label_type, label_expression),
                                           "%s-%s.json" % (filename, idx)))

    def saveFeature(self, filename, idx, transform, feature, label_type, label_expression, writer=None):
        directory_path = self.directory(filename, transform, label_type, label_expression)
        path = os.path.join(directory_path, "%s-%s.json" % (filename, idx))
        row = F.feature_record(feature[0], feature[1], feature[2], frame_no=idx, model_type=self.model_type,
                               transform=transform, filepath=path, label_type=label_type, label_expression=label_expression)
        payload = {'candidate': row['candidate'], 'subset': row['subset'], 'all_hand_peaks': row['all_hand_peaks']}
        if writer is not None:
            writer.put((path, payload))
        else:
            os.makedirs(directory_path, exist_ok=True)
            with open(path, "w") as f:
                json.dump(payload, f)
        return row

    def extract_features_worker(self, video_path, label_type, label_expression, transform='original'):
        filename = os.path.basename(video_path)
        full = os.path.join(self.dataset_base_path, video_path) if self.dataset_base_path else video_path
        os.makedirs(self.directory(filename, transform, label_type, label_expression), exist_ok=True)
        feeder = FrameFeeder(full, self.batch_size, self.rank, self.world_size, limit=self.limit,
                             skip=lambda i: self.is_processed(filename, i, transform, label_type, label_expression))
        writer = queue.Queue()

        def write_loop():
            while True:
                item = writer.get()
                if item is None:
                    return
                with open(item[0], "w") as f:
                    json.dump(item[1], f)
        wt = threading.Thread(target=write_loop, daemon=True)
        wt.start()
        t0 = time.perf_counter()
        order = []

        def batches():
            for idxs, frames in feeder:
                order.append(idxs)
                yield frames, ([self.hand_boxes] * len(idxs) if self.hand_boxes is not None else None)
        rows = []
        ex = self.extractor
        runner = ex.pipeline(batches()) if hasattr(ex.body, "enqueue") else (ex.batch(list(fr.numpy()), hb) for fr, hb in batches())
        rows_s = 0.0
        for bi, res in enumerate(runner):
            t1 = time.perf_counter()
            for idx, (cand, sub, peaks) in zip(order[bi], res):
                rows.append(self.saveFeature(filename, idx, transform, (cand, sub, peaks), label_type, label_expression, writer))
            rows_s += time.perf_counter() - t1
        t1 = time.perf_counter()
        writer.put(None)
        wt.join()
        wall = time.perf_counter() - t0
        # where the wall time went: the reader thread's decode time (overlaps everything else), the consumer's time
        # forming rows (feature_record: Python lists of every candidate), the tail spent waiting for the JSON writer, and
        # the rest = waiting for the GPU pipeline
        self.stats = {"frames": len(rows), "seconds": wall, "frames_per_s": len(rows) / wall if wall > 0 else 0.0,
                      "decode_seconds": feeder.decode_s, "decode_frames_per_s": feeder.frames_decoded / feeder.decode_s
                      if feeder.decode_s > 0 else None, "rows_seconds": rows_s, "writer_tail_seconds": time.perf_counter() - t1,
                      "pipeline_wait_seconds": max(wall - rows_s - (time.perf_counter() - t1), 0.0),
                      "block": (feeder.start, feeder.stop)}
        return rows
