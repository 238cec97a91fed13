"""Hot-path helpers with the reference's names and semantics (src/util.py).

`handDetect` and `npmax` are host scalar math on a handful of numbers per person whose results (integer box corners, the
first arg-max) must be bit-identical to the reference's, and they have one obvious form: they are the reference's own
statements (src/util.py:242-306, 394-399) under the reference's names, comments dropped - a restatement on purpose, not a
rewrite; everything with arithmetic weight on this path (resize, networks, maps, peaks, grouping, hand key points) is CUDA
code of this repository's own design. `padRightDownCorner` and `transfer` keep the reference's signatures."""
import math

import numpy as np


def padRightDownCorner(img, stride, padValue):
    """util.py:12-32 - pad bottom/right with padValue up to a multiple of stride. Returns (img_padded, pad)."""
    h, w = img.shape[0], img.shape[1]
    pad = [0, 0, 0 if h % stride == 0 else stride - h % stride, 0 if w % stride == 0 else stride - w % stride]
    out = np.full((h + pad[2], w + pad[3]) + tuple(img.shape[2:]), padValue, dtype=img.dtype)
    out[:h, :w] = img
    return out, pad


def transfer(model, model_weights):
    """util.py:35-44 - map the flat Caffe-named weight file onto the model's state-dict keys. PoseNet keeps the
    flat names themselves, so for it this is the identity; for an nn.Module the reference's rule applies."""
    keys = list(model.state_dict().keys())
    if all(k in model_weights for k in keys):
        return {k: model_weights[k] for k in keys}
    out = {}
    for name in keys:
        parts = name.split('.')
        out[name] = model_weights['.'.join(parts[3:] if len(parts) > 4 else parts[1:])]
    return out


def handDetect(candidate, subset, oriImg):
    """util.py:242-306 - hand boxes [[x, y, w, is_left], ...] from shoulder/elbow/wrist of every person."""
    ratioWristElbow = 0.33
    detect_result = []
    image_height, image_width = oriImg.shape[0:2]
    for person in np.asarray(subset).astype(int):
        has_left = np.sum(person[[5, 6, 7]] == -1) == 0
        has_right = np.sum(person[[2, 3, 4]] == -1) == 0
        if not (has_left or has_right):
            continue
        hands = []
        if has_left:
            s, e, w = person[[5, 6, 7]]
            hands.append((candidate[s][:2], candidate[e][:2], candidate[w][:2], True))
        if has_right:
            s, e, w = person[[2, 3, 4]]
            hands.append((candidate[s][:2], candidate[e][:2], candidate[w][:2], False))
        for (x1, y1), (x2, y2), (x3, y3), is_left in hands:
            x = x3 + ratioWristElbow * (x3 - x2)
            y = y3 + ratioWristElbow * (y3 - y2)
            distanceWristElbow = math.sqrt((x3 - x2) ** 2 + (y3 - y2) ** 2)
            distanceElbowShoulder = math.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2)
            width = 1.5 * max(distanceWristElbow, 0.9 * distanceElbowShoulder)
            x -= width / 2
            y -= width / 2
            if x < 0:
                x = 0
            if y < 0:
                y = 0
            width1 = width
            width2 = width
            if x + width > image_width:
                width1 = image_width - x
            if y + width > image_height:
                width2 = image_height - y
            width = min(width1, width2)
            if width >= 20:
                detect_result.append([int(x), int(y), int(width), is_left])
    return detect_result


def npmax(array):
    """util.py:394-399 - (row, col) of the first maximum."""
    arrayindex = array.argmax(1)
    arrayvalue = array.max(1)
    i = arrayvalue.argmax()
    j = arrayindex[i]
    return i, j


def gaussian_weights(sigma=3.0, truncate=4.0):
    """The float64 taps scipy.ndimage.gaussian_filter(sigma=3) uses (radius = int(truncate*sigma + 0.5) = 12)."""
    radius = int(truncate * sigma + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


# feature-vector helpers the reference keeps in util (src/util.py:99-151,187-219)
from .features import get_bodypose, get_handpose  # noqa: E402,F401
