"""The sign classifier behind the key points: `translation_model` of demo_isl_translate.py:72-100 and its use
in `ISLSignPosTranslator.call` (src/ISL_Model_parameter.py:322-353) and the demo's rolling window
(demo_isl_translate.py:183-197).

    translator = Translator(weights)            # translation_model.get_weights(), or a file written from it
    probs = translator(window)                  # window [20,156] or [n,20,156] -> float32 [n, classes], on the device
    maxindex, p = translator.top(window)        # np.argmax(encoded_translation), its probability (demo_isl_translate.py:192-194)

The model runs as one CUDA kernel (csrc/translate.cu) through islpose_translate; there is no CPU path. Keras is not needed:
the weights are the arrays `get_weights()` returns, in that order (28 arrays for the reference's architecture), passed as a
list, or saved with `np.savez(path, *translation_model.get_weights())` / `np.save` of their concatenation and given as a
path. The reference's `.keras` archive itself is an HDF5 container; reading it needs h5py, which this image does not have:
export once on the reference side.
"""
import ctypes as C

import numpy as np

from . import _lib
from .features import N_FEATURES, WINDOW

UNITS = 32


def weight_shapes(n_classes, n_features=N_FEATURES, units=UNITS, hidden=32):
    """Shapes of `translation_model.get_weights()` (demo_isl_translate.py:72-99), in order."""
    s = [(n_features,)] * 4
    for fin in (n_features, 2 * units):
        s += [(fin, 4 * units), (units, 4 * units), (4 * units,)] * 2
    s += [(2 * units, hidden)] + [(hidden,)] * 4 + [(hidden, hidden)] + [(hidden,)] * 4 + [(hidden, n_classes), (n_classes,)]
    return s


def random_weights(n_classes, seed=0):
    """Seeded stand-in for `translation_model.get_weights()` (the trained model does not ship with the reference): uniform
    kernels, BatchNorm statistics in the range of pixel coordinates. For benchmarks and smoke runs."""
    rng = np.random.RandomState(seed)
    out = []
    for i, shp in enumerate(weight_shapes(n_classes)):
        if len(shp) == 2:
            lim = np.sqrt(6.0 / (shp[0] + shp[1]))
            w = rng.uniform(-lim, lim, shp)
        elif i == 2:
            w = rng.uniform(0.0, 300.0, shp)        # moving_mean of the input normalisation
        elif i == 3:
            w = rng.uniform(2000.0, 20000.0, shp)   # moving_variance of the input normalisation
        else:
            w = rng.uniform(0.2, 1.2, shp)
        out.append(w.astype(np.float32))
    return out


def load_weights(source):
    """list of arrays | .npz written by np.savez(path, *get_weights()) | .npy of the concatenation -> list of float32 arrays
    or one flat float32 array."""
    if isinstance(source, (list, tuple)):
        return [np.asarray(a, dtype=np.float32) for a in source]
    if isinstance(source, np.ndarray):
        return np.ascontiguousarray(source, dtype=np.float32).reshape(-1)
    path = str(source)
    if path.endswith(".npz"):
        with np.load(path) as z:
            keys = sorted(z.files, key=lambda k: int(k.split("_")[-1]))   # arr_0, arr_1, ...
            return [z[k].astype(np.float32) for k in keys]
    return np.load(path).astype(np.float32).reshape(-1)


class Translator(object):
    def __init__(self, weights, n_classes=None, device=None):
        import torch

        L = _lib.lib()
        if not torch.cuda.is_available():
            raise _lib.IslposeError("Translator needs a CUDA device (there is no CPU fallback)")
        w = load_weights(weights)
        if isinstance(w, list):
            if len(w) != 28:
                raise ValueError("the reference's classifier has 28 weight arrays (demo_isl_translate.py:72-99), got %d" % len(w))
            n_classes = int(w[-1].shape[0])
            for i, (a, shp) in enumerate(zip(w, weight_shapes(n_classes))):
                if tuple(a.shape) != tuple(shp):
                    raise ValueError("weight %d has shape %s, expected %s" % (i, tuple(a.shape), tuple(shp)))
            flat = np.concatenate([a.reshape(-1) for a in w])
        else:
            flat = w
            if n_classes is None:   # the class count follows from the length: 33 floats per class behind a fixed prefix
                fixed = int(L.islpose_translate_weight_floats(1)) - 33
                if (flat.size - fixed) % 33 != 0:
                    raise ValueError("a flat weight array of %d floats fits no class count" % flat.size)
                n_classes = (flat.size - fixed) // 33
        if flat.size != int(L.islpose_translate_weight_floats(n_classes)):
            raise ValueError("%d classes need %d weight floats, got %d" % (n_classes, L.islpose_translate_weight_floats(n_classes), flat.size))
        self.n_classes = int(n_classes)
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._w = torch.from_numpy(np.ascontiguousarray(flat)).to(self.device)

    def __call__(self, window):
        """window: numpy or device tensor, [T,156] or [n,T,156] (T <= 32; all-zero rows are masked steps). Returns the class
        probabilities as a float32 device tensor [n, classes] (`translation_layer(...)`, ISL_Model_parameter.py:353)."""
        import torch

        if not torch.is_tensor(window):
            window = torch.from_numpy(np.ascontiguousarray(np.asarray(window, dtype=np.float64)))
        w = window.to(self.device, dtype=torch.float64).contiguous()
        if w.dim() == 2:
            w = w.unsqueeze(0)
        if w.dim() != 3 or w.shape[2] != N_FEATURES:
            raise ValueError("windows are [n, T, %d], got %s" % (N_FEATURES, tuple(w.shape)))
        n, T = int(w.shape[0]), int(w.shape[1])
        probs = torch.empty((n, self.n_classes), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().islpose_translate(_lib.ptr(w), n, T, N_FEATURES, _lib.ptr(self._w), C.c_int64(self._w.numel()),
                                                    self.n_classes, _lib.ptr(probs), _lib.stream_ptr()), "islpose_translate")
        return probs

    def sliding(self, rows, length=WINDOW):
        """rows: [T,156] feature rows of a clip (numpy or device tensor). Classifies every window of `length` consecutive
        rows the demo loop would form (demo_isl_translate.py:183-192: rows i-19..i for i >= 20) in one launch; returns the
        float32 device tensor [max(T - length, 0), classes]."""
        import torch

        if not torch.is_tensor(rows):
            rows = torch.from_numpy(np.ascontiguousarray(np.asarray(rows, dtype=np.float64)))
        r = rows.to(self.device, dtype=torch.float64)
        n = int(r.shape[0]) - length
        if n <= 0:
            return torch.empty((0, self.n_classes), dtype=torch.float32, device=self.device)
        wins = r.unfold(0, length, 1)[1:].permute(0, 2, 1).contiguous()   # window k = rows k+1 .. k+length
        return self(wins)

    def top(self, window):
        """(class index, probability) per window: demo_isl_translate.py:192-194."""
        p = self(window).cpu().numpy()
        idx = p.argmax(axis=1)
        return idx, p[np.arange(p.shape[0]), idx]


class RollingTranslator(object):
    """The demo's loop (demo_isl_translate.py:183-197): the last 20 feature rows, classified once the window is full."""

    def __init__(self, translator, length=WINDOW):
        self.translator = translator
        self.length = length
        self.rows = []

    def push(self, feature_row):
        """feature_row: float64 [156] (KeypointExtractor.features / features.frame_features). Returns None while the window fills
        (the reference classifies from frame `length` + 1 on), then (class index, probability) of the current window."""
        row = np.asarray(feature_row, dtype=np.float64).reshape(N_FEATURES)
        if len(self.rows) < self.length:     # demo_isl_translate.py:183-184: the first `length` frames only fill the window
            self.rows.append(row)
            return None
        self.rows = self.rows[1:] + [row]    # :186-187
        idx, p = self.translator.top(np.stack(self.rows))
        return int(idx[0]), float(p[0])
