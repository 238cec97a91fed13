"""Feature-vector formation after the key-point path: the step that turns one frame's
(candidate, subset, all_hand_peaks) into the 156 numbers the ISL classifier consumes, and the 20-frame window.

Reference: util.get_bodypose (src/util.py:99-151), util.get_handpose (src/util.py:187-219), populate_features
(ISL_model_xy.py:78-112 = src/ISL_Model_parameter.py:376-410) and the sliding window (ISL_Model_parameter.py:370-374).
Same names, same argument meaning, same return structures. These are the HOST restatements: the product forms the 156
numbers on the device (csrc/features.cu through KeypointExtractor.features / pipeline(with_features=True)) and the tests use
the functions here as the checker; `feature_record` / `feature_json` build the reference's per-frame row and JSON payload
(lists of Python numbers, by nature host work).

Layout of the 156-vector: 15 body circle x, 15 body circle y (circles in joint-major, person-minor order, missing
ones 0), then per hand (2 hands): 21 x, 21 y, 21 key-point indices ("peak text" 0..20 as float).

Documented divergence: util.get_handpose holds two result slots, so a third hand raises IndexError in the
reference (util.py:198,205); here hands beyond the second are ignored.
"""
import math

import numpy as np

from .tables import model_dims

LIMB_SEQ_BODY25 = [[1, 0], [1, 2], [2, 3], [3, 4], [1, 5], [5, 6], [6, 7], [1, 8], [8, 9], [9, 10], [10, 11], [8, 12],
                   [12, 13], [13, 14], [0, 15], [0, 16], [15, 17], [16, 18], [11, 24], [11, 22], [14, 21], [14, 19],
                   [22, 23], [19, 20]]
LIMB_SEQ_COCO = [[1, 2], [1, 5], [2, 3], [3, 4], [5, 6], [6, 7], [1, 8], [8, 9], [9, 10], [1, 11], [11, 12], [12, 13],
                 [1, 0], [0, 14], [14, 16], [0, 15], [15, 17], [2, 16], [5, 17]]
HAND_EDGES = [[0, 1], [1, 2], [2, 3], [3, 4], [0, 5], [5, 6], [6, 7], [7, 8], [0, 9], [9, 10], [10, 11], [11, 12],
              [0, 13], [13, 14], [14, 15], [15, 16], [0, 17], [17, 18], [18, 19], [19, 20]]
N_BODY_CIRCLES = 15
N_FEATURES = 2 * N_BODY_CIRCLES + 2 * 3 * 21   # 156
WINDOW = 20


def get_bodypose(candidate, subset, model_type='coco'):
    """-> (x_y_circles, x_y_sticks): joint positions (joint-major, person-minor, util.py:123-130) and per limb
    (mean x, mean y, angle in degrees, length) for limbs with both ends present (util.py:133-145)."""
    limb_seq = LIMB_SEQ_BODY25 if model_type == 'body25' else LIMB_SEQ_COCO
    njoint = model_dims('body25' if model_type == 'body25' else 'coco')[0] - 1
    circles = []
    for i in range(njoint):
        for n in range(len(subset)):
            index = int(subset[n][i])
            if index == -1:
                continue
            x, y = candidate[index][0:2]
            circles.append((x, y))
    sticks = []
    for i in range(njoint - 1):   # util.py:133 walks njoint-1 limbs (17 of coco's 19, 24 of body25's 24)
        for n in range(len(subset)):
            index = subset[n][np.array(limb_seq[i])]
            if -1 in index:
                continue
            Y = candidate[index.astype(int), 0]
            X = candidate[index.astype(int), 1]
            mX = np.mean(X)
            mY = np.mean(Y)
            length = ((X[0] - X[1]) ** 2 + (Y[0] - Y[1]) ** 2) ** 0.5
            angle = math.degrees(math.atan2(X[0] - X[1], Y[0] - Y[1]))
            sticks.append((mY, mX, angle, length))
    return (circles, sticks)


def get_handpose(all_hand_peaks, show_number=False):
    """-> (export_edges, export_peaks), two slots each (util.py:187-219): per hand the edges whose two ends were
    found, as (edge index, (x1, y1), (x2, y2)), and all 21 key points as (x, y, str(index))."""
    export_edges = [[], []]
    export_peaks = [[], []]
    for idx, peaks in enumerate(all_hand_peaks[:2]):
        peaks = np.asarray(peaks)
        for ie, e in enumerate(HAND_EDGES):
            if np.sum(np.all(peaks[e], axis=1) == 0) == 0:
                x1, y1 = peaks[e[0]]
                x2, y2 = peaks[e[1]]
                export_edges[idx].append((ie, (x1, y1), (x2, y2)))
        for i, keypoint in enumerate(peaks):
            x, y = keypoint
            export_peaks[idx].append((x, y, str(i)))
    return (export_edges, export_peaks)


def populate_features(bodypose_circles, handpose_peaks):
    """-> float array [156] (ISL_model_xy.py:78-112)."""
    feature = []
    for col in (0, 1):
        for idx in range(N_BODY_CIRCLES):
            feature.append(bodypose_circles[idx][col] if idx < len(bodypose_circles) else 0)
    for hand_idx in range(2):
        for col in (0, 1, 2):
            for idx in range(21):
                feature.append(float(handpose_peaks[hand_idx][idx][col]) if idx < len(handpose_peaks[hand_idx]) else 0)
    return np.array(feature)


def frame_features(candidate, subset, all_hand_peaks, model_type='coco'):
    """One frame's 156-vector from the extractor's outputs (the chain of demo_isl_translate.py's frame loop)."""
    circles, _ = get_bodypose(candidate, subset, model_type)
    _, peaks = get_handpose(all_hand_peaks)
    return populate_features(circles, peaks).astype(np.float64)


def feature_record(candidate, subset, all_hand_peaks, frame_no=0, model_type='body25', transform='original',
                   filepath=None, label_type=None, label_expression=None):
    """One frame's row of the feature-extraction loop (extract_features.py:105-141 `saveFeature`, the dict that
    `saveFeaturesDict` turns into a DataFrame / CSV row): the three extractor outputs as lists plus the derived circles,
    sticks, hand edges and hand key points. Same keys, same value structures; nothing is written to disk here."""
    circles, sticks = get_bodypose(candidate, subset, model_type)
    edges, peaks = get_handpose(all_hand_peaks)
    return {
        'transform': transform,
        'filepath': filepath,
        'frame_no': frame_no,
        'type': label_type,
        'expression': label_expression,
        'candidate': np.asarray(candidate).tolist(),
        'subset': np.asarray(subset).tolist(),
        'all_hand_peaks': [np.asarray(p).tolist() for p in all_hand_peaks],
        'bodypose_x_ytupple': circles,
        'bodypose_x_y_sticks': sticks,
        'handpose_edges': edges,
        'handpose_peaks': peaks,
    }


def feature_json(candidate, subset, all_hand_peaks):
    """The per-frame JSON payload the reference writes next to every frame (extract_features.py:112-117)."""
    import json

    return json.dumps({'candidate': np.asarray(candidate).tolist(), 'subset': np.asarray(subset).tolist(),
                       'all_hand_peaks': [np.asarray(p).tolist() for p in all_hand_peaks]})


class FeatureWindow(object):
    """The classifier's input window: the last WINDOW frames, oldest first (ISL_Model_parameter.py:370-374)."""

    def __init__(self, length=WINDOW, n_features=N_FEATURES):
        self.window = np.zeros((length, n_features), dtype=np.float64)
        self.count = 0

    def push(self, feature):
        self.window[:-1] = self.window[1:]
        self.window[-1] = feature
        self.count += 1
        return self.window

    @property
    def full(self):
        return self.count >= self.window.shape[0]
