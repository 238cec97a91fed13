"""The frame loop around the key-point path (isl_b200/frames.py) with stand-in estimators: block sharding,
the pinned-buffer feeder, resume by existence, the per-frame JSON files and feature rows of extract_features.py:105-173."""
import json
import os

import numpy as np
import pytest

import isl_b200  # noqa: F401
from isl_b200 import features as F
from isl_b200 import frames as FR
from isl_b200.extract import KeypointExtractor


class FakeBody:
    model_type = "coco"

    def batch(self, frames):
        out = []
        for f in frames:
            s = int(np.asarray(f, dtype=np.int64).sum())
            n = s % 5
            out.append((np.arange(n * 4, dtype=np.float64).reshape(n, 4) + s % 1000 if n else np.array([]), -1 * np.ones((0, 20))))
        return out


def _clip(n, h=48, w=64):
    return np.stack([np.random.RandomState(i).randint(0, 256, (h, w, 3)).astype(np.uint8) for i in range(n)])


def test_block_ranges_partition_in_order():
    for n in (0, 1, 7, 30, 31):
        for world in (1, 2, 4, 8):
            spans = [FR.block_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1


@pytest.mark.parametrize("world", [1, 3])
def test_feeder_delivers_this_ranks_block_in_batches(world):
    clip = _clip(17)
    seen = {}
    for rank in range(world):
        feeder = FR.FrameFeeder(clip, batch_size=4, rank=rank, world_size=world)
        for idxs, batch in feeder:
            assert tuple(batch.shape[1:]) == (48, 64, 3) and len(idxs) == batch.shape[0] <= 4
            for i, fr in zip(idxs, batch.numpy()):
                assert i not in seen
                seen[i] = fr.copy()   # copied at once: the feeder recycles a buffer two batches later
    assert sorted(seen) == list(range(17))
    assert all(np.array_equal(seen[i], clip[i]) for i in range(17))


def test_feeder_skips_without_decoding_and_reads_video_files(tmp_path):
    cv2 = pytest.importorskip("cv2")
    path = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 30, (64, 48))
    if not wr.isOpened():
        pytest.skip("this OpenCV build cannot write MJPG")
    base = cv2.resize(np.random.RandomState(0).randint(0, 256, (6, 8, 3)).astype(np.uint8), (64, 48), interpolation=cv2.INTER_CUBIC)
    for i in range(12):
        wr.write(np.roll(base, 4 * i, axis=1))
    wr.release()
    ref = []
    cap = cv2.VideoCapture(path)
    while True:
        ok, fr = cap.read()
        if not ok:
            break
        ref.append(fr)
    assert len(ref) == 12
    feeder = FR.FrameFeeder(path, batch_size=5, skip=lambda i: i % 3 == 0)
    got = {}
    for idxs, batch in feeder:
        for i, fr in zip(idxs, batch.numpy()):
            got[i] = fr.copy()
    assert sorted(got) == [i for i in range(12) if i % 3]
    assert all(np.array_equal(got[i], ref[i]) for i in got)   # skipping by grab() keeps the decoder in step
    assert feeder.frames_decoded == 8


def test_video_extractor_writes_the_reference_files_and_resumes(tmp_path):
    clip = _clip(10)
    np.save(str(tmp_path / "MVI_0001.npy"), clip)
    ex = KeypointExtractor(FakeBody(), None)
    vx = FR.VideoExtractor(ex, str(tmp_path / "transforms"), dataset_base_path=str(tmp_path), batch_size=3)
    rows = vx.extract_features_worker("MVI_0001.npy", "Adjectives", "loud")
    assert [r["frame_no"] for r in rows] == list(range(10))
    d = tmp_path / "transforms" / "Adjectives" / "loud" / "MVI_0001-original"
    want = FakeBody().batch(list(clip))
    for i, r in enumerate(rows):
        assert list(r) == list(F.feature_record(want[i][0], want[i][1], [], frame_no=i))   # saveFeature's keys, in order
        assert r["filepath"] == str(d / ("MVI_0001.npy-%d.json" % i)) and r["type"] == "Adjectives" and r["expression"] == "loud"
        payload = json.load(open(r["filepath"]))
        assert sorted(payload) == ["all_hand_peaks", "candidate", "subset"]
        assert payload["candidate"] == np.asarray(want[i][0]).tolist()
        assert vx.is_processed("MVI_0001.npy", i, "original", "Adjectives", "loud")
    # resume by existence (extract_features.py:97-101, 157-159): only frames without a JSON are processed again
    for i in (2, 7):
        os.remove(str(d / ("MVI_0001.npy-%d.json" % i)))
    again = vx.extract_features_worker("MVI_0001.npy", "Adjectives", "loud")
    assert [r["frame_no"] for r in again] == [2, 7] and vx.stats["frames"] == 2
    # two ranks: contiguous blocks that merge into the whole clip
    vx2 = [FR.VideoExtractor(ex, str(tmp_path / "t2"), dataset_base_path=str(tmp_path), batch_size=4, rank=r, world_size=2)
           for r in range(2)]
    parts = [v.extract_features_worker("MVI_0001.npy", "a", "b") for v in vx2]
    assert [r["frame_no"] for p in parts for r in p] == list(range(10))
