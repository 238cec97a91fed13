"""Feature-vector step against golden vectors produced by the unmodified reference functions
(tests/golden/make_golden_features.py). Exact equality: the step is index and float64 copy work plus atan2 / sqrt
on a handful of values, evaluated by the same numpy / math calls as the reference."""
import glob
import os

import numpy as np
import pytest

import isl_b200  # noqa: F401
from isl_b200 import features as F
from isl_b200.extract import KeypointExtractor

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
G = np.load(os.path.join(GOLD, "features.npz"))
CASES = sorted(set(k.split("/")[0] for k in G.files))


@pytest.mark.parametrize("case", CASES)
def test_features_match_reference(case):
    d = np.load(os.path.join(GOLD, "body_%s.npz" % case), allow_pickle=True)
    mt = str(d["model_type"])
    hands = [h for h in G[case + "/hands"]]
    circles, sticks = F.get_bodypose(d["candidate"], d["subset"], mt)
    edges, peaks = F.get_handpose(hands)
    assert np.array_equal(np.array(circles, dtype=np.float64).reshape(-1, 2), G[case + "/circles"])
    assert np.array_equal(np.array(sticks, dtype=np.float64).reshape(-1, 4), G[case + "/sticks"])
    assert [len(e) for e in edges] == G[case + "/n_edges"].tolist()
    feat = F.populate_features(circles, peaks)
    assert feat.shape == (156,) and np.array_equal(feat.astype(np.float64), G[case + "/feature"])
    assert np.array_equal(F.frame_features(d["candidate"], d["subset"], hands, mt), G[case + "/feature"])


def test_empty_frame_and_extra_hands():
    f = F.frame_features(np.array([]), -1 * np.ones((0, 20)), [], "coco")
    assert f.shape == (156,) and not f.any()
    hands = [np.full((21, 2), k + 1, dtype=np.int64) for k in range(3)]   # the reference raises on a third hand
    f = F.frame_features(np.array([]), -1 * np.ones((0, 20)), hands, "coco")
    assert f[30] == 1 and f[30 + 63] == 2 and f[30 + 42:30 + 63].tolist() == list(range(21))


def test_window_slides_oldest_first():
    w = F.FeatureWindow()
    for t in range(25):
        out = w.push(np.full(156, float(t)))
    assert w.full and out.shape == (20, 156) and out[0, 0] == 5.0 and out[-1, 0] == 24.0


def test_clip_features_with_stand_in_estimators():
    """KeypointExtractor.features on estimators without a device path: order and shape of the clip matrix."""
    d = np.load(os.path.join(GOLD, "body_coco_p3_s1.npz"), allow_pickle=True)

    class Body(object):
        model_type = "coco"

        def batch(self, frames):
            return [(d["candidate"], d["subset"]) if int(f[0, 0, 0]) % 2 == 0 else (np.array([]), -1 * np.ones((0, 20)))
                    for f in frames]

    frames = [np.full((int(d["h"]), int(d["w"]), 3), t, dtype=np.uint8) for t in range(5)]
    X = KeypointExtractor(Body(), None).features(frames, batch_size=2)
    assert X.shape == (5, 156)
    want = F.frame_features(d["candidate"], d["subset"], [], "coco")
    assert np.array_equal(X[0], want) and np.array_equal(X[2], want) and not X[1].any()


def test_feature_record_has_the_reference_row_structure():
    """extract_features.py:105-141: keys and value structures of the per-frame row, and the JSON payload of :112-117."""
    import json

    d = np.load(os.path.join(GOLD, "body_body25_p5_s2.npz"), allow_pickle=True)
    hands = [h for h in G["coco_p2_s4/hands"]]
    row = F.feature_record(d["candidate"], d["subset"], hands, frame_no=7, model_type="body25", label_type="train",
                           label_expression="hello")
    assert list(row) == ['transform', 'filepath', 'frame_no', 'type', 'expression', 'candidate', 'subset', 'all_hand_peaks',
                         'bodypose_x_ytupple', 'bodypose_x_y_sticks', 'handpose_edges', 'handpose_peaks']
    assert row['frame_no'] == 7 and row['transform'] == 'original' and row['type'] == 'train'
    assert row['candidate'] == d["candidate"].tolist() and row['subset'] == d["subset"].tolist()
    assert np.array_equal(np.array(row['bodypose_x_ytupple']).reshape(-1, 2), G["body25_p5_s2/circles"])
    assert np.array_equal(np.array(row['bodypose_x_y_sticks']).reshape(-1, 4), G["body25_p5_s2/sticks"])
    assert len(row['handpose_peaks']) == 2 and len(row['handpose_peaks'][0]) == 21 and row['handpose_peaks'][0][3][2] == '3'
    payload = json.loads(F.feature_json(d["candidate"], d["subset"], hands))
    assert sorted(payload) == ['all_hand_peaks', 'candidate', 'subset'] and payload['subset'] == d["subset"].tolist()
    json.dumps(row['candidate'])   # every numeric list is JSON-serialisable, as the reference needs for its CSV / JSON
