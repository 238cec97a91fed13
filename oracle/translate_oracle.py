"""TEST INFRASTRUCTURE - CPU restatement (numpy, float32) of the reference's sign classifier, the Keras `Sequential` built in
/root/reference/demo_isl_translate.py:72-100 and applied to a 20 x 156 window in
/root/reference/src/ISL_Model_parameter.py:322-353 (`ISLSignPosTranslator.call`).

PARITY UNPINNED against Keras: `keras` is a third-party dependency of the reference (requirements.txt: keras 3, torch
backend, `KERAS_BACKEND=torch` at demo_isl_translate.py:17) that is absent from this image, and the reference ships neither
the trained weights (`model/isl_model_final.keras`) nor outputs of this model. What is restated here is Keras 3's published
inference algorithm for each layer, each function citing the reference line that instantiates the layer:

* Masking(mask_value=0.)           a time step is masked when every feature equals 0 (keras/src/layers/core/masking.py)
* BatchNormalization()             (x - moving_mean) / sqrt(moving_var + 1e-3) * gamma + beta, the mask passes through
* Bidirectional(LSTM(32, ...))     gate order i, f, c, o; sigmoid recurrent activation, tanh activation; at a masked step the
                                   states are kept and the output repeats the previous output (zeros before the first
                                   unmasked step) - keras/src/backend/torch/rnn.py `rnn`; the backward layer scans the flipped
                                   sequence and, with return_sequences, its outputs are flipped back (bidirectional.py)
* Dropout                          identity at inference; recurrent_dropout likewise
* Activation('elu'), Dense(use_bias=False), Dense(softmax)

The one part an independent implementation in this image can pin is pinned: the LSTM layers (equations, gate order,
bidirectional concatenation, trailing-mask final states) are checked against torch.nn.LSTM / packed sequences in
tests/test_translate.py. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import numpy as np

WINDOW, N_FEATURES, UNITS, HIDDEN = 20, 156, 32, 32
BN_EPS = np.float32(1e-3)  # keras BatchNormalization default epsilon


def weight_shapes(n_classes, n_features=N_FEATURES, units=UNITS, hidden=HIDDEN):
    """`translation_model.get_weights()` order of demo_isl_translate.py:72-99 (name, shape)."""
    s = []
    s += [("bn0/" + k, (n_features,)) for k in ("gamma", "beta", "moving_mean", "moving_variance")]
    for layer, fin in (("lstm1", n_features), ("lstm2", 2 * units)):
        for d in ("forward", "backward"):
            s += [("%s/%s/kernel" % (layer, d), (fin, 4 * units)), ("%s/%s/recurrent_kernel" % (layer, d), (units, 4 * units)),
                  ("%s/%s/bias" % (layer, d), (4 * units,))]
    s += [("dense1/kernel", (2 * units, hidden))]
    s += [("bn1/" + k, (hidden,)) for k in ("gamma", "beta", "moving_mean", "moving_variance")]
    s += [("dense2/kernel", (hidden, hidden))]
    s += [("bn2/" + k, (hidden,)) for k in ("gamma", "beta", "moving_mean", "moving_variance")]
    s += [("dense3/kernel", (hidden, n_classes)), ("dense3/bias", (n_classes,))]
    return s


def make_weights(n_classes, seed=0, n_features=N_FEATURES):
    """Seeded stand-in for the trained weights (none ship with the reference): Keras' default initialisers in spirit -
    glorot-uniform kernels, orthogonal-free small recurrent kernels, forget-gate bias 1, he-normal dense kernels - and
    non-trivial BatchNorm statistics so that every term of the normalisation is exercised."""
    rng = np.random.RandomState(seed)
    out = []
    for name, shape in weight_shapes(n_classes, n_features):
        leaf = name.rsplit("/", 1)[1]
        if leaf == "gamma":
            w = rng.uniform(0.5, 1.5, shape)
        elif leaf == "beta":
            w = rng.uniform(-0.3, 0.3, shape)
        elif leaf == "moving_mean":
            w = rng.uniform(0.0, 300.0, shape) if name.startswith("bn0") else rng.uniform(-0.5, 0.5, shape)
        elif leaf == "moving_variance":
            w = rng.uniform(2000.0, 20000.0, shape) if name.startswith("bn0") else rng.uniform(0.2, 2.0, shape)
        elif leaf == "bias" and "lstm" in name:
            w = np.zeros(shape)
            w[UNITS:2 * UNITS] = 1.0  # unit_forget_bias
            w = w + rng.uniform(-0.05, 0.05, shape)
        elif leaf == "bias":
            w = rng.uniform(-0.1, 0.1, shape)
        else:
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            w = rng.uniform(-lim, lim, shape)
        out.append(w.astype(np.float32))
    return out


def _sigmoid(x):
    return (np.float32(1) / (np.float32(1) + np.exp(-x))).astype(np.float32)


def _elu(x):
    return np.where(x > 0, x, np.exp(np.minimum(x, 0)) - np.float32(1)).astype(np.float32)


def _bn(x, gamma, beta, mean, var):
    return ((x - mean) / np.sqrt(var + BN_EPS) * gamma + beta).astype(np.float32)


def lstm_scan(x, mask, kernel, recurrent, bias, go_backwards=False):
    """keras LSTM(units) over x [T, F] with a boolean step mask [T]; returns the output sequence [T, units] in the order
    the layer scanned (demo_isl_translate.py:76,81: LSTM(32, recurrent_dropout=0.2[, return_sequences=True]))."""
    T = x.shape[0]
    u = recurrent.shape[0]
    order = range(T - 1, -1, -1) if go_backwards else range(T)
    h = np.zeros((u,), np.float32)
    c = np.zeros((u,), np.float32)
    prev = np.zeros((u,), np.float32)
    outs = []
    for t in order:
        if mask[t]:
            z = (x[t] @ kernel + h @ recurrent + bias).astype(np.float32)
            i, f, g, o = _sigmoid(z[:u]), _sigmoid(z[u:2 * u]), np.tanh(z[2 * u:3 * u]).astype(np.float32), _sigmoid(z[3 * u:])
            c = (f * c + i * g).astype(np.float32)
            h = (o * np.tanh(c)).astype(np.float32)
            prev = h
        outs.append(prev)
    return np.stack(outs)


def bilstm(x, mask, w6, return_sequences):
    """keras Bidirectional(LSTM(...)) with merge_mode='concat' (demo_isl_translate.py:76,81)."""
    fwd = lstm_scan(x, mask, w6[0], w6[1], w6[2], go_backwards=False)
    bwd = lstm_scan(x, mask, w6[3], w6[4], w6[5], go_backwards=True)
    if return_sequences:
        return np.concatenate([fwd, bwd[::-1]], axis=1)
    return np.concatenate([fwd[-1], bwd[-1]])


def translate(window, weights):
    """window [T, F] (any float type; ISL_Model_parameter.py:353 reshapes the 20 rows to (1, 20, 156)) -> softmax over the
    classes, float32 [n_classes]."""
    w = [np.asarray(a, np.float32) for a in weights]
    x = np.asarray(window).astype(np.float32)
    mask = np.any(x != 0, axis=-1)                      # Masking(mask_value=0.), demo_isl_translate.py:74
    x = _bn(x, *w[0:4])                                 # BatchNormalization(), :75
    x = bilstm(x, mask, w[4:10], return_sequences=True)   # :76
    x = bilstm(x, mask, w[10:16], return_sequences=False)  # :81 (Dropout :80 is the identity at inference)
    x = _elu(x)                                         # :83
    x = (x @ w[16]).astype(np.float32)                  # Dense(32, use_bias=False), :84
    x = _elu(_bn(x, *w[17:21]))                         # :89-91
    x = (x @ w[21]).astype(np.float32)                  # :92
    x = _elu(_bn(x, *w[22:26]))                         # :96-97
    z = (x @ w[26] + w[27]).astype(np.float32)          # Dense(n_classes, softmax), :99
    e = np.exp(z - z.max())
    return (e / e.sum()).astype(np.float32)


def translate_batch(windows, weights):
    return np.stack([translate(wd, weights) for wd in windows])
