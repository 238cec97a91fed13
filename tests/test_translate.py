"""The sign classifier (demo_isl_translate.py:72-99). Keras is absent from the image, so the oracle restates
Keras' published layer algorithms ("parity unpinned" against Keras itself); what an independent implementation here can
pin is pinned: the LSTM layers against torch.nn.LSTM. The CUDA kernel is then compared with the oracle."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import isl_b200  # noqa: E402,F401
from isl_b200 import translate as T  # noqa: E402
from oracle import translate_oracle as O  # noqa: E402

N_CLASSES = 167  # len(expression_mapping) of the reference (src/expression_mapping.py: keys 0..166)


def _windows(n, seed, steps=20, trailing_blank=0, holes=()):
    """Feature-like rows: pixel coordinates, zeros for joints that were not found, blank rows where asked."""
    rng = np.random.RandomState(seed)
    w = rng.uniform(0, 720, (n, steps, 156))
    w[rng.uniform(size=w.shape) < 0.15] = 0.0
    if trailing_blank:
        w[:, steps - trailing_blank:, :] = 0.0
    for h in holes:
        w[:, h, :] = 0.0
    return w


def _torch_bilstm(w6, fin):
    m = torch.nn.LSTM(fin, 32, batch_first=True, bidirectional=True)
    with torch.no_grad():
        for sfx, (k, r, b) in (("", w6[0:3]), ("_reverse", w6[3:6])):
            getattr(m, "weight_ih_l0" + sfx).copy_(torch.from_numpy(k.T.copy()))   # same gate order: i, f, g(c), o
            getattr(m, "weight_hh_l0" + sfx).copy_(torch.from_numpy(r.T.copy()))
            getattr(m, "bias_ih_l0" + sfx).copy_(torch.from_numpy(b))
            getattr(m, "bias_hh_l0" + sfx).zero_()
    return m.eval()


def test_oracle_lstm_layers_equal_torch_lstm():
    """Unmasked: the whole bidirectional sequence output; trailing blank steps: the final states (packed sequences)."""
    w = O.make_weights(N_CLASSES, seed=3)
    rng = np.random.RandomState(0)
    x = rng.randn(20, 156).astype(np.float32)
    mask = np.ones(20, bool)
    m = _torch_bilstm(w[4:10], 156)
    with torch.no_grad():
        ref, (hn, _) = m(torch.from_numpy(x)[None])
    got = O.bilstm(x, mask, w[4:10], return_sequences=True)
    assert np.abs(got - ref[0].numpy()).max() < 2e-6
    last = O.bilstm(x, mask, w[4:10], return_sequences=False)
    assert np.abs(last - np.concatenate([hn[0, 0].numpy(), hn[1, 0].numpy()])).max() < 2e-6
    # 13 live steps followed by 7 masked ones: keras keeps the states through masked steps = torch's packed sequence of length 13
    mask[13:] = False
    packed = torch.nn.utils.rnn.pack_padded_sequence(torch.from_numpy(x)[None], [13], batch_first=True)
    with torch.no_grad():
        out_p, (hn, _) = m(packed)
    seq13 = torch.nn.utils.rnn.pad_packed_sequence(out_p, batch_first=True)[0][0].numpy()
    last = O.bilstm(x, mask, w[4:10], return_sequences=False)
    assert np.abs(last - np.concatenate([hn[0, 0].numpy(), hn[1, 0].numpy()])).max() < 2e-6
    got = O.bilstm(x, mask, w[4:10], return_sequences=True)
    assert np.abs(got[:13] - seq13).max() < 2e-6
    # masked steps repeat the previous output: forward = step 12's, backward (scanned first, flipped back) = zeros
    assert np.array_equal(got[13:, :32], np.repeat(got[12:13, :32], 7, axis=0)) and not got[13:, 32:].any()


def test_oracle_outputs_are_distributions_and_mask_matters():
    w = O.make_weights(N_CLASSES, seed=1)
    full = _windows(1, 5)[0]
    p = O.translate(full, w)
    assert p.shape == (N_CLASSES,) and abs(float(p.sum()) - 1.0) < 1e-5 and (p >= 0).all()
    holed = full.copy()
    holed[7] = 0.0
    assert np.abs(O.translate(holed, w) - p).max() > 1e-6
    blank = np.zeros((20, 156))          # every step masked: both LSTMs return zeros, the head still yields a distribution
    assert abs(float(O.translate(blank, w).sum()) - 1.0) < 1e-5


def test_weight_shapes_agree_with_the_oracle():
    assert [s for _, s in O.weight_shapes(N_CLASSES)] == T.weight_shapes(N_CLASSES)
    assert len(T.weight_shapes(N_CLASSES)) == 28


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["full", "trailing_blank", "holes", "short"])
def test_translate_kernel_equals_the_oracle(case, tmp_path):
    w = O.make_weights(N_CLASSES, seed=2)
    if case == "full":
        win = _windows(5, 11)
    elif case == "trailing_blank":
        win = _windows(3, 12, trailing_blank=6)
    elif case == "holes":
        win = _windows(3, 13, holes=(0, 4, 5, 19))
        win[2] = 0.0                      # one window entirely blank
    else:
        win = _windows(2, 14, steps=7)
    tr = T.Translator(w)
    probs = tr(win).cpu().numpy()
    ref = O.translate_batch(win, w)
    assert probs.shape == ref.shape
    assert np.abs(probs - ref).max() < 2e-5, np.abs(probs - ref).max()       # float32 both sides; summation order differs
    assert np.array_equal(probs.argmax(1), ref.argmax(1))
    # the same weights through a file, and a 2-D window
    path = str(tmp_path / "w.npz")
    np.savez(path, *w)
    tr2 = T.Translator(path)
    assert tr2.n_classes == N_CLASSES
    assert np.array_equal(tr2(win[0]).cpu().numpy(), probs[:1])
    idx, p = tr.top(win)
    assert np.array_equal(idx, probs.argmax(1)) and np.allclose(p, probs.max(1))


@pytest.mark.gpu
def test_rolling_translator_follows_the_demo_loop():
    """demo_isl_translate.py:183-197: nothing for the first 20 frames, then one classification per frame on the last 20 rows."""
    w = O.make_weights(N_CLASSES, seed=4)
    tr = T.Translator(w)
    rows = _windows(1, 21, steps=25)[0]
    roll = T.RollingTranslator(tr)
    got = [roll.push(r) for r in rows]
    assert all(g is None for g in got[:20]) and all(g is not None for g in got[20:])
    for i in range(20, 25):
        ref = O.translate(rows[i - 19:i + 1], w)
        assert got[i][0] == int(ref.argmax()) and abs(got[i][1] - float(ref.max())) < 2e-5
    # the same windows in one launch
    p = tr.sliding(rows).cpu().numpy()
    assert p.shape == (5, N_CLASSES)
    assert [int(i) for i in p.argmax(1)] == [g[0] for g in got[20:]]
    assert tr.sliding(rows[:20]).shape[0] == 0


@pytest.mark.gpu
def test_translate_abi_error_paths():
    import ctypes as C
    from isl_b200 import _lib
    L = _lib.lib()
    need = L.islpose_translate_weight_floats(N_CLASSES)
    assert need == sum(int(np.prod(s)) for s in T.weight_shapes(N_CLASSES))
    wdev = torch.zeros(need, dtype=torch.float32, device="cuda")
    win = torch.zeros((1, 20, 156), dtype=torch.float64, device="cuda")
    out = torch.zeros((1, N_CLASSES), dtype=torch.float32, device="cuda")
    args = lambda T_=20, F=156, nw=need, cls=N_CLASSES: (_lib.ptr(win), 1, T_, F, _lib.ptr(wdev), C.c_int64(nw), cls, _lib.ptr(out), _lib.stream_ptr())
    assert L.islpose_translate(*args()) == 0
    assert L.islpose_translate(*args(T_=33)) != 0 and b"window length" in L.islpose_last_error()
    assert L.islpose_translate(*args(F=150)) != 0 and b"156" in L.islpose_last_error()
    assert L.islpose_translate(*args(nw=need - 1)) != 0 and b"weight floats" in L.islpose_last_error()
    with pytest.raises(ValueError):
        T.Translator(O.make_weights(N_CLASSES)[:-1])
