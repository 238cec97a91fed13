"""CPU oracle for the OpenPose keypoint-extraction hot path (body + hand).

TEST INFRASTRUCTURE - NOT PRODUCT CODE. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module, and only as the checker. The product package never does.

It restates, in numpy / torch-CPU fp32, what the reference computes on this path:
    src/model.py   (three VGG-19 + CPM networks)           -> net_forward()
    src/body.py    (Body.__call__)                          -> body_maps(), body_peaks(), body_connections(),
                                                               body_assemble(), body_call()
    src/hand.py    (Hand.__call__)                          -> hand_maps(), hand_peaks(), hand_call()
    src/util.py    (padRightDownCorner, handDetect, npmax)  -> pad_right_down_corner(), hand_detect(), npmax()
Each function cites the reference lines it follows.

Third-party arithmetic the reference calls but does not contain (all unpinned by the reference,
requirements.txt:1-6; versions are the ones in this image):
    opencv-python 4.13.0  cv2.resize(INTER_CUBIC)           -> resize_cubic()   (restated; see below)
    scipy 1.18.1          ndimage.gaussian_filter(sigma=3)  -> gaussian_filter_sigma3() (restated)
    scikit-image (absent) measure.label(connectivity=2)     -> label8()        (restated)
    torch 2.11.0          conv2d / max_pool2d / prelu       -> called directly (fp32 CPU = the float reference)
`backend="lib"` makes the oracle call cv2 / scipy exactly where the reference does (same speed as the
reference; used for the CPU baseline), `backend="restated"` uses the numpy restatements. tests/ checks that
the two agree bit for bit wherever OpenCV runs its own code; the one exception is documented at resize_cubic().

PARITY PINNING: the reference ships no tests or golden vectors for this path, so the
oracle is pinned against outputs of the reference itself, run in the build container through
oracle/ref_import.py; the vectors are committed under tests/golden/ with the script that made them
(tests/golden/make_golden.py) and tests/test_oracle_golden.py replays them on every run.
"""
import math

import numpy as np

f32 = np.float32

# ----------------------------------------------------------------------------------------------------------
# cv2.resize(..., interpolation=cv2.INTER_CUBIC)  -- OpenCV 4.13 modules/imgproc/src/resize.cpp, generic path
# ----------------------------------------------------------------------------------------------------------


def _cubic_coeffs(frac):
    """interpolateCubic(): Keys cubic with A = -0.75, evaluated in float32 in OpenCV's operation order."""
    a = f32(-0.75)
    x = frac.astype(f32)
    one = f32(1)
    x1 = x + one
    c0 = ((a * x1 - f32(5) * a) * x1 + f32(8) * a) * x1 - f32(4) * a
    c1 = ((a + f32(2)) * x - (a + f32(3))) * x * x + one
    xm = one - x
    c2 = ((a + f32(2)) * xm - (a + f32(3))) * xm * xm + one
    c3 = one - c0 - c1 - c2
    return np.stack([c0, c1, c2, c3], -1).astype(f32)


def _cubic_tables(ssize, dsize, scale):
    """Per destination index: the 4 clamped source indices (replicate border, applied per tap) and the
    4 float32 weights. src = (dst + 0.5) * scale - 0.5 is evaluated in double and rounded to float."""
    d = np.arange(dsize, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(f32)
    s = np.floor(f).astype(np.int64)
    frac = (f - s.astype(f32)).astype(f32)
    idx = np.clip(s[:, None] + np.arange(-1, 3)[None, :], 0, ssize - 1)
    return idx, _cubic_coeffs(frac)


def resize_out_size(h, w, fx, fy):
    """dsize when the caller passes fx/fy: saturate_cast<int>(src * f) = round half to even."""
    return int(np.rint(h * fy)), int(np.rint(w * fx))


def resize_cubic(src, dsize=None, fx=None, fy=None):
    """cv2.resize(src, dsize or (0,0), fx, fy, INTER_CUBIC) for HxWxC uint8 or float32 arrays.

    float32: horizontal pass ((S0*a0 + S1*a1) + S2*a2) + S3*a3, vertical pass in the order of OpenCV's
    4-lane SIMD body ((S3*b3 + S2*b2) + S1*b1) + S0*b0 with the last (W*C mod 4) elements of each row in the
    scalar order ((S0*b0 + S1*b1) + S2*b2) + S3*b3; no fused multiply-add (SSE3 baseline build). Bit-exact
    against cv2 for the channel counts on this path (19, 22, 26, 38, 52).
    uint8: 11-bit fixed-point weights (cvRound(w * 2048)), integer horizontal pass, vertical pass in float
    (weights * 2^-22, SIMD order) with round-half-even and saturation. Bit-exact against cv2 with
    cv2.ipp.setUseIPP(False). The pip wheel dispatches 1/3/4-channel uint8 and float images to closed-source
    Intel IPP instead; against that the result differs by at most 1 grey level in about 5 % of the pixels
    (measured in tests/test_oracle_primitives.py). Only the network input (body.py:53, hand.py:37) is affected.
    """
    h, w = src.shape[:2]
    if dsize is None:
        dh, dw = resize_out_size(h, w, fx, fy)
        sx, sy = 1.0 / fx, 1.0 / fy
    else:
        dw, dh = dsize
        sx, sy = 1.0 / (dw / float(w)), 1.0 / (dh / float(h))
    xi, xa = _cubic_tables(w, dw, sx)
    yi, ya = _cubic_tables(h, dh, sy)
    squeeze = src.ndim == 2
    s3 = src[:, :, None] if squeeze else src
    c = s3.shape[2]
    if src.dtype == np.uint8:
        ia = np.rint(xa * f32(2048)).astype(np.int64)
        ib = np.rint(ya * f32(2048)).astype(np.int64)
        s64 = s3.astype(np.int64)
        hz = np.zeros((h, dw, c), np.int64)
        for j in range(4):
            hz += s64[:, xi[:, j], :] * ia[None, :, j, None]
        bf = (ib.astype(f32) * (f32(1.0) / f32(2048 * 2048))).astype(f32)
        hf = hz.astype(f32)
        v = None
        for k in (3, 2, 1, 0):
            t = hf[yi[:, k]] * bf[:, k, None, None]
            v = t if v is None else v + t
        out = np.clip(np.rint(v), 0, 255).astype(np.uint8)
    elif src.dtype == np.float32:
        hz = None
        for j in range(4):
            t = s3[:, xi[:, j], :] * xa[None, :, j, None]
            hz = t if hz is None else hz + t
        rows = hz.reshape(h, dw * c)  # OpenCV's vertical pass sees a row as W*C floats

        def vertical(order, sl):
            v = None
            for k in order:
                t = rows[yi[:, k]][:, sl] * ya[:, k, None]
                v = t if v is None else v + t
            return v

        body = (dw * c) // 4 * 4
        out = np.empty((dh, dw * c), f32)
        out[:, :body] = vertical((3, 2, 1, 0), slice(0, body))
        if body < dw * c:
            out[:, body:] = vertical((0, 1, 2, 3), slice(body, dw * c))
        out = out.reshape(dh, dw, c)
    else:
        raise TypeError("resize_cubic: uint8 or float32 only, got %s" % src.dtype)
    return out[:, :, 0] if squeeze else out


def _resize(src, backend, dsize=None, fx=None, fy=None):
    if backend == "lib":
        import cv2

        if dsize is None:
            return cv2.resize(src, (0, 0), fx=fx, fy=fy, interpolation=cv2.INTER_CUBIC)
        return cv2.resize(src, dsize, interpolation=cv2.INTER_CUBIC)
    return resize_cubic(src, dsize=dsize, fx=fx, fy=fy)


# ----------------------------------------------------------------------------------------------------------
# scipy.ndimage.gaussian_filter(plane, sigma=3)  -- scipy 1.18 ndimage/_filters.py + src/ni_filters.c
# ----------------------------------------------------------------------------------------------------------

GAUSS_RADIUS = 12  # int(truncate * sigma + 0.5) with truncate = 4.0, sigma = 3


def gaussian_weights_sigma3():
    x = np.arange(-GAUSS_RADIUS, GAUSS_RADIUS + 1)
    phi = np.exp(-0.5 / (3.0 * 3.0) * x.astype(np.float64) ** 2)
    return phi / phi.sum()


def _reflect_index(n, r):
    """mode='reflect' (d c b a | a b c d | d c b a): extended index -> source index, for any n >= 1."""
    i = np.arange(-r, n + r)
    period = 2 * n
    i = np.mod(i, period)
    return np.where(i >= n, period - 1 - i, i)


def _correlate1d_symmetric(a, w, axis):
    """NI_Correlate1D, symmetric branch: tmp = x[l]*w[0]; for jj = -r..-1: tmp += (x[l+jj] + x[l-jj]) * w[jj]."""
    a = np.moveaxis(a, axis, 0)
    n = a.shape[0]
    r = (len(w) - 1) // 2
    ext = a[_reflect_index(n, r)]
    acc = ext[r:r + n] * w[r]
    for jj in range(-r, 0):
        acc = acc + (ext[r + jj:r + jj + n] + ext[r - jj:r - jj + n]) * w[r + jj]
    return np.moveaxis(acc, 0, axis)


def gaussian_filter_sigma3(plane):
    """float64 separable 25-tap Gaussian, axis 0 then axis 1, reflect borders (body.py:88, hand.py:61)."""
    w = gaussian_weights_sigma3()
    a = np.asarray(plane, dtype=np.float64)
    return _correlate1d_symmetric(_correlate1d_symmetric(a, w, 0), w, 1)


def _gaussian(plane, backend):
    if backend == "lib":
        from scipy.ndimage import gaussian_filter

        return gaussian_filter(plane, sigma=3)
    return gaussian_filter_sigma3(plane)


# ----------------------------------------------------------------------------------------------------------
# skimage.measure.label(binary, connectivity=2) -- 8-connected components, labels in raster order of first pixel
# ----------------------------------------------------------------------------------------------------------


def label8(binary):
    """Two-pass union-find labelling. Returns (labels int32 HxW with 0 = background, count)."""
    b = np.asarray(binary) != 0
    h, w = b.shape
    lab = np.zeros((h, w), np.int32)
    parent = [0]

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    for y in range(h):
        row = b[y]
        xs = np.nonzero(row)[0]
        for x in xs:
            best = 0
            for (yy, xx) in ((y, x - 1), (y - 1, x - 1), (y - 1, x), (y - 1, x + 1)):
                if yy >= 0 and 0 <= xx < w and lab[yy, xx]:
                    r = find(lab[yy, xx])
                    if best == 0:
                        best = r
                    elif r != best:
                        lo, hi = (best, r) if best < r else (r, best)
                        parent[hi] = lo
                        best = lo
            if best == 0:
                parent.append(len(parent))
                best = len(parent) - 1
            lab[y, x] = best
    roots = np.array([find(i) for i in range(len(parent))], np.int32)
    # renumber roots in raster order of each component's first pixel (= ascending provisional label of the root)
    uniq = np.unique(roots[1:]) if len(parent) > 1 else np.array([], np.int32)
    remap = np.zeros(len(parent), np.int32)
    for new, r in enumerate(uniq, start=1):
        remap[r] = new
    final = remap[roots]
    return final[lab], len(uniq)


def _label(binary, backend):
    if backend == "lib":
        from scipy import ndimage

        return ndimage.label(binary, structure=np.ones((3, 3), np.int32))
    return label8(binary)


# ----------------------------------------------------------------------------------------------------------
# src/util.py
# ----------------------------------------------------------------------------------------------------------


def pad_right_down_corner(img, stride=8, pad_value=128):
    """util.py:12-32: constant pad on the bottom/right up to a multiple of `stride`. Returns (padded, pad)
    with pad = [up, left, down, right] like the reference."""
    h, w = img.shape[:2]
    down = 0 if h % stride == 0 else stride - h % stride
    right = 0 if w % stride == 0 else stride - w % stride
    out = np.full((h + down, w + right) + img.shape[2:], pad_value, dtype=img.dtype)
    out[:h, :w] = img
    return out, [0, 0, down, right]


def npmax(a):
    """util.py:394-399: (row, col) of the first row holding the global maximum, first column in that row."""
    cols = a.argmax(1)
    vals = a.max(1)
    i = int(vals.argmax())
    return i, int(cols[i])


def hand_detect(candidate, subset, image_shape):
    """util.py:242-306. image_shape = oriImg.shape; returns [[x, y, w, is_left], ...] (ints, bool)."""
    ratio = 0.33
    image_height, image_width = image_shape[0:2]
    out = []
    for person in np.asarray(subset).astype(int):
        has_left = not np.any(person[[5, 6, 7]] == -1)
        has_right = not np.any(person[[2, 3, 4]] == -1)
        hands = []
        if has_left:
            hands.append((person[[5, 6, 7]], True))
        if has_right:
            hands.append((person[[2, 3, 4]], False))
        for (si, ei, wi), is_left in hands:
            x1, y1 = candidate[si][:2]
            x2, y2 = candidate[ei][:2]
            x3, y3 = candidate[wi][:2]
            x = x3 + ratio * (x3 - x2)
            y = y3 + ratio * (y3 - y2)
            d_we = math.sqrt((x3 - x2) ** 2 + (y3 - y2) ** 2)
            d_es = math.sqrt((x2 - x1) ** 2 + (y2 - y1) ** 2)
            width = 1.5 * max(d_we, 0.9 * d_es)
            x -= width / 2
            y -= width / 2
            if x < 0:
                x = 0
            if y < 0:
                y = 0
            w1 = width
            w2 = width
            if x + width > image_width:
                w1 = image_width - x
            if y + width > image_height:
                w2 = image_height - y
            width = min(w1, w2)
            if width >= 20:
                out.append([int(x), int(y), int(width), is_left])
    return out


# ----------------------------------------------------------------------------------------------------------
# src/model.py -- the three networks, as tables over flat Caffe layer names (the on-disk weight format,
# util.py:35-44) and a functional fp32 forward.
# ----------------------------------------------------------------------------------------------------------

_VGG_PREFIX = [("conv1_1", 3, 64), ("conv1_2", 64, 64), "pool", ("conv2_1", 64, 128), ("conv2_2", 128, 128), "pool",
               ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), ("conv3_4", 256, 256), "pool",
               ("conv4_1", 256, 512), ("conv4_2", 512, 512)]


def net_layers(kind):
    """Every conv of the network as (name, cin, cout, k, act, prelu_name) with act in {'relu','prelu','none'},
    in a deterministic order (model.py:66-207 body25, :210-329 coco, :331-407 hand)."""
    out = []

    def add(name, cin, cout, k, act, prelu=None):
        out.append((name, cin, cout, k, act, prelu))

    if kind == "coco":
        for item in _VGG_PREFIX:
            if item != "pool":
                add(item[0], item[1], item[2], 3, "relu")
        add("conv4_3_CPM", 512, 256, 3, "relu")
        add("conv4_4_CPM", 256, 128, 3, "relu")
        for br, cout in ((1, 38), (2, 19)):
            for i in (1, 2, 3):
                add("conv5_%d_CPM_L%d" % (i, br), 128, 128, 3, "relu")
            add("conv5_4_CPM_L%d" % br, 128, 512, 1, "relu")
            add("conv5_5_CPM_L%d" % br, 512, cout, 1, "none")
        for st in range(2, 7):
            for br, cout in ((1, 38), (2, 19)):
                add("Mconv1_stage%d_L%d" % (st, br), 185, 128, 7, "relu")
                for i in (2, 3, 4, 5):
                    add("Mconv%d_stage%d_L%d" % (i, st, br), 128, 128, 7, "relu")
                add("Mconv6_stage%d_L%d" % (st, br), 128, 128, 1, "relu")
                # model.py:215-218 lists 'Mconv7_stage6_L1' twice and omits 'Mconv7_stage6_L2': the final
                # heat-map layer keeps its ReLU.
                last_act = "relu" if (st == 6 and br == 2) else "none"
                add("Mconv7_stage%d_L%d" % (st, br), 128, cout, 1, last_act)
    elif kind == "hand":
        for item in _VGG_PREFIX:
            if item != "pool":
                add(item[0], item[1], item[2], 3, "relu")
        for name in ("conv4_3", "conv4_4", "conv5_1", "conv5_2"):
            add(name, 512, 512, 3, "relu")
        add("conv5_3_CPM", 512, 128, 3, "relu")
        add("conv6_1_CPM", 128, 512, 1, "relu")
        add("conv6_2_CPM", 512, 22, 1, "none")
        for st in range(2, 7):
            add("Mconv1_stage%d" % st, 150, 128, 7, "relu")
            for i in (2, 3, 4, 5):
                add("Mconv%d_stage%d" % (i, st), 128, 128, 7, "relu")
            add("Mconv6_stage%d" % st, 128, 128, 1, "relu")
            add("Mconv7_stage%d" % st, 128, 22, 1, "none")
    elif kind == "body25":
        for item in _VGG_PREFIX:
            if item != "pool":
                name = item[0]
                if name == "conv4_2":
                    add(name, item[1], item[2], 3, "prelu", "prelu4_2")
                else:
                    add(name, item[1], item[2], 3, "relu")
        add("conv4_3_CPM", 512, 256, 3, "prelu", "prelu4_3_CPM")
        add("conv4_4_CPM", 256, 128, 3, "prelu", "prelu4_4_CPM")
        for (br, stage, cin0, width, mid, cout) in _BODY25_STAGES:
            tag = "stage%d_L%d" % (stage, br)
            for blk in range(1, 6):
                cin = cin0 if blk == 1 else 3 * width
                for j in range(3):
                    add("Mconv%d_%s_%d" % (blk, tag, j), cin if j == 0 else width, width, 3, "prelu",
                        "Mprelu%d_%s_%d" % (blk, tag, j))
            add("Mconv6_%s" % tag, 3 * width, mid, 1, "prelu", "Mprelu6_%s" % tag)
            add("Mconv7_%s" % tag, mid, cout, 1, "none")
    else:
        raise ValueError(kind)
    return out


# (branch, stage, first-block Cin, block width, 1x1 width, outputs) in execution order (model.py:179-207)
_BODY25_STAGES = [(2, 0, 128, 96, 256, 52), (2, 1, 180, 128, 512, 52), (2, 2, 180, 128, 512, 52),
                  (2, 3, 180, 128, 512, 52), (1, 0, 180, 96, 256, 26), (1, 1, 206, 128, 512, 26)]


def make_flat_weights(kind, seed=0, gain=1.0, head_gain=1.0, init="he"):
    """Seeded random weights in the reference's on-disk format: a flat dict of Caffe layer names
    ('conv1_1.weight', 'Mprelu1_stage0_L2_0.weight', ...) -> float32 tensors (util.py:35-44 maps these onto
    the module's state-dict keys).
      init="he":    He-uniform scaled by `gain` (keeps activations alive through ~30 ReLU layers, so the maps cross
                    the thresholds and peak finding / grouping get exercised); `head_gain` scales the stage outputs.
      init="torch": the distribution nn.Conv2d / nn.PReLU are constructed with (kaiming_uniform(a=sqrt(5)):
                    U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weight and bias, PReLU slope 0.25) - the "random-init
                    weights" of SURVEY.md section 8d; produces (almost) no peaks, like the reference does untrained."""
    import torch

    g = torch.Generator().manual_seed(seed)
    w = {}
    for (name, cin, cout, k, act, prelu) in net_layers(kind):
        if init == "torch":
            bound = 1.0 / math.sqrt(cin * k * k)
            bias_bound = bound
        else:
            bound = gain * math.sqrt(6.0 / (cin * k * k))
            bias_bound = 0.05
            if name.startswith("Mconv7") or name in ("conv5_5_CPM_L1", "conv5_5_CPM_L2", "conv6_2_CPM"):
                bound *= head_gain
        w[name + ".weight"] = (torch.rand((cout, cin, k, k), generator=g) * 2 - 1) * bound
        w[name + ".bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bias_bound
        if prelu is not None:
            w[prelu + ".weight"] = torch.full((cout,), 0.25) if init == "torch" else torch.rand((cout,), generator=g) * 0.3
    return w


def net_forward(kind, weights, x, emulate_bf16=False, device=None):
    """fp32 forward (model.py), on CPU unless `device` says otherwise. x: torch.float32 [N,3,h,w]. Returns (PAF, heat)
    for bodies, heat for the hand.

    emulate_bf16: a numerical model of the CUDA path, for tests that want to separate its two error sources - bf16
    operands (modelled here: weights, the network input and every stored activation rounded to bf16, fp32 accumulation,
    bias / activation in fp32, the last stage's outputs left in fp32) from accumulation order (not modelled)."""
    import torch
    import torch.nn.functional as F

    spec = {l[0]: l for l in net_layers(kind)}
    if device is not None:
        weights = {k: v.to(device) for k, v in weights.items()}
        x = x.to(device)
    rnd = (lambda t: t.to(torch.bfloat16).to(torch.float32)) if emulate_bf16 else (lambda t: t)
    if emulate_bf16:
        weights = {k: (rnd(v) if v.ndim == 4 else v) for k, v in weights.items()}
        x = rnd(x)
    last = {"coco": ("Mconv7_stage6_L1", "Mconv7_stage6_L2"), "hand": ("Mconv7_stage6",),
            "body25": ("Mconv7_stage3_L2", "Mconv7_stage1_L1")}[kind]
    f32_outs = {}

    def conv(name, t):
        _, _, _, k, act, prelu = spec[name]
        t = F.conv2d(t, weights[name + ".weight"], weights[name + ".bias"], padding=(k - 1) // 2)
        if act == "relu":
            t = F.relu(t)
        elif act == "prelu":
            t = F.prelu(t, weights[prelu + ".weight"])
        if name in last:
            f32_outs[name] = t   # the network heads leave in fp32; what later stages read of them is the bf16 copy
        return rnd(t)

    def backbone(t, names):
        for n in names:
            t = F.max_pool2d(t, 2, 2) if n == "pool" else conv(n, t)
        return t

    prefix = [i if i == "pool" else i[0] for i in _VGG_PREFIX]
    with torch.no_grad():
        if kind == "coco":
            feat = backbone(x, prefix + ["conv4_3_CPM", "conv4_4_CPM"])
            outs = []
            for br in (1, 2):
                outs.append(backbone(feat, ["conv5_%d_CPM_L%d" % (i, br) for i in (1, 2, 3, 4, 5)]))
            for st in range(2, 7):
                cat = torch.cat([outs[0], outs[1], feat], 1)  # model.py:308-324
                outs = [backbone(cat, ["Mconv%d_stage%d_L%d" % (i, st, br) for i in range(1, 8)]) for br in (1, 2)]
            return f32_outs[last[0]], f32_outs[last[1]]
        if kind == "hand":
            feat = backbone(x, prefix + ["conv4_3", "conv4_4", "conv5_1", "conv5_2", "conv5_3_CPM"])
            out = backbone(feat, ["conv6_1_CPM", "conv6_2_CPM"])
            for st in range(2, 7):
                out = backbone(torch.cat([out, feat], 1), ["Mconv%d_stage%d" % (i, st) for i in range(1, 8)])
            return f32_outs[last[0]]
        if kind == "body25":
            feat = backbone(x, prefix + ["conv4_3_CPM", "conv4_4_CPM"])

            def stage(t, br, st):
                tag = "stage%d_L%d" % (st, br)
                for blk in range(1, 6):  # model.py:171-177: three chained 3x3 convs, outputs concatenated
                    parts = []
                    for j in range(3):
                        t = conv("Mconv%d_%s_%d" % (blk, tag, j), t)
                        parts.append(t)
                    t = torch.cat(parts, 1)
                return conv("Mconv7_%s" % tag, conv("Mconv6_%s" % tag, t))

            t = feat
            paf = None
            for st in range(4):
                paf = stage(t, 2, st)
                t = torch.cat([feat, paf], 1)
            heat0 = stage(t, 1, 0)
            stage(torch.cat([feat, heat0, paf], 1), 1, 1)
            return f32_outs[last[0]], f32_outs[last[1]]
    raise ValueError(kind)


# ----------------------------------------------------------------------------------------------------------
# src/body.py
# ----------------------------------------------------------------------------------------------------------

LIMB_SEQ = {
    "coco": [[1, 2], [1, 5], [2, 3], [3, 4], [5, 6], [6, 7], [1, 8], [8, 9], [9, 10], [1, 11], [11, 12], [12, 13],
             [1, 0], [0, 14], [14, 16], [0, 15], [15, 17], [2, 16], [5, 17]],
    "body25": [[1, 0], [1, 2], [2, 3], [3, 4], [1, 5], [5, 6], [6, 7], [1, 8], [8, 9], [9, 10], [10, 11], [8, 12],
               [12, 13], [13, 14], [0, 15], [0, 16], [15, 17], [16, 18], [11, 24], [11, 22], [14, 21], [14, 19],
               [22, 23], [19, 20]],
}
MAP_IDX = {
    "coco": [[12, 13], [20, 21], [14, 15], [16, 17], [22, 23], [24, 25], [0, 1], [2, 3], [4, 5], [6, 7], [8, 9],
             [10, 11], [28, 29], [30, 31], [34, 35], [32, 33], [36, 37], [18, 19], [26, 27]],
    "body25": [[30, 31], [14, 15], [16, 17], [18, 19], [22, 23], [24, 25], [26, 27], [0, 1], [6, 7], [2, 3], [4, 5],
               [8, 9], [10, 11], [12, 13], [32, 33], [34, 35], [36, 37], [38, 39], [50, 51], [46, 47], [44, 45],
               [40, 41], [48, 49], [42, 43]],
}


def model_dims(model_type):
    return (26, 52) if model_type == "body25" else (19, 38)


def preprocess(img, scale, backend="restated", stride=8, pad_value=128):
    """body.py:53-56 / hand.py:37-40: cubic resize by `scale`, pad to a multiple of 8 with 128, x/256 - 0.5,
    HWC -> 1CHW float32. Returns (data [1,3,h,w] float32 numpy, padded_shape, pad)."""
    small = _resize(np.ascontiguousarray(img), backend, fx=scale, fy=scale)
    padded, pad = pad_right_down_corner(small, stride, pad_value)
    im = np.transpose(np.float32(padded[:, :, :, np.newaxis]), (3, 2, 0, 1)) / 256 - 0.5
    return np.ascontiguousarray(im), padded.shape, pad


def upsample_to_image(lowres_chw, padded_shape, pad, out_hw, backend="restated", stride=8):
    """body.py:69-72,75-78 / hand.py:51-54: CHW -> HWC, x8 cubic, crop the padding off, cubic to (W, H)."""
    m = np.transpose(np.asarray(lowres_chw, dtype=f32), (1, 2, 0))
    m = _resize(np.ascontiguousarray(m), backend, fx=stride, fy=stride)
    m = m[:padded_shape[0] - pad[2], :padded_shape[1] - pad[3], :]
    return _resize(np.ascontiguousarray(m), backend, dsize=(out_hw[1], out_hw[0]))


def body_maps(net_fn, img, model_type="coco", scale_search=(0.5,), boxsize=368, backend="restated"):
    """body.py:47-81. net_fn(data float32 [1,3,h,w] numpy) -> (paf [C,h/8,w/8], heat [C,h/8,w/8]) numpy.
    Returns float64 (heatmap_avg [H,W,njoint], paf_avg [H,W,npaf]).
    Keeps quirk Q1 (body.py:80): `heatmap_avg += heatmap_avg + heatmap / S` doubles the running sum, so with
    S scales the weights are 2^(S-1-s)/S rather than 1/S; paf_avg (body.py:81) is a true mean."""
    njoint, npaf = model_dims(model_type)
    h, w = img.shape[:2]
    heat_avg = np.zeros((h, w, njoint))
    paf_avg = np.zeros((h, w, npaf))
    n = len(scale_search)
    for s in scale_search:
        scale = s * boxsize / h
        data, pshape, pad = preprocess(img, scale, backend)
        paf, heat = net_fn(data)
        heat_full = upsample_to_image(heat, pshape, pad, (h, w), backend)
        paf_full = upsample_to_image(paf, pshape, pad, (h, w), backend)
        heat_avg += heat_avg + heat_full / n
        paf_avg += +paf_full / n
    return heat_avg, paf_avg


def body_peaks(heat_avg, njoint, thre1=0.1, backend="restated"):
    """body.py:83-107. Returns all_peaks: per part a list of (x, y, score, id); x,y numpy int64 like the
    reference. 4-neighbour test against zero-filled shifts (quirk Q8), score from the unsmoothed map."""
    all_peaks = []
    counter = 0
    for part in range(njoint - 1):
        map_ori = heat_avg[:, :, part]
        sm = _gaussian(map_ori, backend)
        up = np.zeros(sm.shape)
        up[1:, :] = sm[:-1, :]
        down = np.zeros(sm.shape)
        down[:-1, :] = sm[1:, :]
        left = np.zeros(sm.shape)
        left[:, 1:] = sm[:, :-1]
        right = np.zeros(sm.shape)
        right[:, :-1] = sm[:, 1:]
        binary = (sm >= up) & (sm >= down) & (sm >= left) & (sm >= right) & (sm > thre1)
        ys, xs = np.nonzero(binary)  # row-major order
        peaks = [(xs[i], ys[i], map_ori[ys[i], xs[i]], counter + i) for i in range(len(xs))]
        all_peaks.append(peaks)
        counter += len(xs)
    return all_peaks


def body_connections(all_peaks, paf_avg, model_type, image_height, thre2=0.05, mid_num=10):
    """body.py:128-178. Returns (connection_all, special_k)."""
    limb_seq, map_idx = LIMB_SEQ[model_type], MAP_IDX[model_type]
    connection_all = []
    special_k = []
    for k in range(len(map_idx)):
        score_mid = paf_avg[:, :, map_idx[k]]
        cand_a = all_peaks[limb_seq[k][0]]
        cand_b = all_peaks[limb_seq[k][1]]
        na, nb = len(cand_a), len(cand_b)
        if na == 0 or nb == 0:
            special_k.append(k)
            connection_all.append([])
            continue
        cands = []
        for i in range(na):
            for j in range(nb):
                vec = np.subtract(cand_b[j][:2], cand_a[i][:2])
                norm = math.sqrt(vec[0] * vec[0] + vec[1] * vec[1])
                norm = max(0.001, norm)
                vec = np.divide(vec, norm)
                xs = np.linspace(cand_a[i][0], cand_b[j][0], num=mid_num)
                ys = np.linspace(cand_a[i][1], cand_b[j][1], num=mid_num)
                vx = np.array([score_mid[int(round(ys[t])), int(round(xs[t])), 0] for t in range(mid_num)])
                vy = np.array([score_mid[int(round(ys[t])), int(round(xs[t])), 1] for t in range(mid_num)])
                mid = np.multiply(vx, vec[0]) + np.multiply(vy, vec[1])
                prior = sum(mid) / len(mid) + min(0.5 * image_height / norm - 1, 0)
                ok1 = len(np.nonzero(mid > thre2)[0]) > 0.8 * len(mid)
                if ok1 and prior > 0:
                    cands.append([i, j, prior, prior + cand_a[i][2] + cand_b[j][2]])
        cands = sorted(cands, key=lambda c: c[2], reverse=True)  # stable
        conn = np.zeros((0, 5))
        for (i, j, s, _) in cands:
            if i not in conn[:, 3] and j not in conn[:, 4]:
                conn = np.vstack([conn, [cand_a[i][3], cand_b[j][3], s, i, j]])
                if len(conn) >= min(na, nb):
                    break
        connection_all.append(conn)
    return connection_all, special_k


def body_assemble(all_peaks, connection_all, special_k, model_type, strict=True):
    """body.py:180-235. Returns (candidate, subset). strict=True raises IndexError on a third matching row
    exactly like the reference's two-slot subset_idx (quirk Q5); strict=False keeps the first two matches,
    which is what the CUDA path does."""
    njoint, _ = model_dims(model_type)
    limb_seq = LIMB_SEQ[model_type]
    subset = -1 * np.ones((0, njoint + 1))
    candidate = np.array([item for sub in all_peaks for item in sub])
    for k in range(len(limb_seq)):
        if k in special_k:
            continue
        conn = connection_all[k]
        part_as = conn[:, 0]
        part_bs = conn[:, 1]
        ia, ib = limb_seq[k]
        for i in range(len(conn)):
            found = 0
            idx = [-1, -1]
            for j in range(len(subset)):
                if subset[j][ia] == part_as[i] or subset[j][ib] == part_bs[i]:
                    if found >= 2:
                        if strict:
                            raise IndexError("list assignment index out of range")
                        continue
                    idx[found] = j
                    found += 1
            if found == 1:
                j = idx[0]
                if subset[j][ib] != part_bs[i]:
                    subset[j][ib] = part_bs[i]
                    subset[j][-1] += 1
                    subset[j][-2] += candidate[part_bs[i].astype(int), 2] + conn[i][2]
            elif found == 2:
                j1, j2 = idx
                membership = ((subset[j1] >= 0).astype(int) + (subset[j2] >= 0).astype(int))[:-2]
                if len(np.nonzero(membership == 2)[0]) == 0:
                    subset[j1][:-2] += (subset[j2][:-2] + 1)
                    subset[j1][-2:] += subset[j2][-2:]
                    subset[j1][-2] += conn[i][2]
                    subset = np.delete(subset, j2, 0)
                else:
                    subset[j1][ib] = part_bs[i]
                    subset[j1][-1] += 1
                    subset[j1][-2] += candidate[part_bs[i].astype(int), 2] + conn[i][2]
            elif not found and k < njoint - 2:
                row = -1 * np.ones(njoint + 1)
                row[ia] = part_as[i]
                row[ib] = part_bs[i]
                row[-1] = 2
                row[-2] = sum(candidate[conn[i, :2].astype(int), 2]) + conn[i][2]
                subset = np.vstack([subset, row])
    drop = [i for i in range(len(subset)) if subset[i][-1] < 4 or subset[i][-2] / subset[i][-1] < 0.4]
    subset = np.delete(subset, drop, axis=0)
    return candidate, subset


def body_from_maps(heat_avg, paf_avg, model_type, thre1=0.1, thre2=0.05, mid_num=10, backend="restated",
                   strict=True):
    njoint, _ = model_dims(model_type)
    peaks = body_peaks(heat_avg, njoint, thre1, backend)
    conn, special = body_connections(peaks, paf_avg, model_type, heat_avg.shape[0], thre2, mid_num)
    return body_assemble(peaks, conn, special, model_type, strict)


def body_call(net_fn, img, model_type="coco", scale_search=(0.5,), backend="restated", strict=True):
    """Body.__call__ (body.py:39-235) -> (candidate, subset)."""
    heat_avg, paf_avg = body_maps(net_fn, img, model_type, scale_search, backend=backend)
    return body_from_maps(heat_avg, paf_avg, model_type, backend=backend, strict=strict)


# ----------------------------------------------------------------------------------------------------------
# src/hand.py
# ----------------------------------------------------------------------------------------------------------


def hand_maps(net_fn, img, scale_search=(0.5, 1.0, 1.5, 2.0), boxsize=368, backend="restated"):
    """hand.py:31-56 -> float64 heatmap_avg [h,w,22] (a true mean over the scales)."""
    h, w = img.shape[:2]
    avg = np.zeros((h, w, 22))
    for s in scale_search:
        scale = s * boxsize / h
        data, pshape, pad = preprocess(img, scale, backend)
        heat = net_fn(data)
        avg += upsample_to_image(heat, pshape, pad, (h, w), backend) / len(scale_search)
    return avg


def hand_peaks(heat_avg, thre=0.05, backend="restated"):
    """hand.py:58-74 -> int array [21,2] of (x, y); [0,0] = not found. heat_avg is modified in place like
    the reference does through its view (quirk Q7)."""
    out = []
    for part in range(21):
        map_ori = heat_avg[:, :, part]
        sm = _gaussian(map_ori, backend)
        binary = np.ascontiguousarray(sm > thre, dtype=np.uint8)
        if np.sum(binary) == 0:
            out.append([0, 0])
            continue
        lab, num = _label(binary, backend)
        mass = [np.sum(map_ori[lab == i]) for i in range(1, num + 1)]
        keep = int(np.argmax(mass)) + 1
        map_ori[lab != keep] = 0
        y, x = npmax(map_ori)
        out.append([x, y])
    return np.array(out)


def hand_call(net_fn, img, backend="restated"):
    """Hand.__call__ (hand.py:24-74)."""
    return hand_peaks(hand_maps(net_fn, img, backend=backend), backend=backend)


# ----------------------------------------------------------------------------------------------------------
# convenience: a torch-CPU net_fn over flat weights
# ----------------------------------------------------------------------------------------------------------


def make_net_fn(kind, weights, emulate_bf16=False):
    import torch

    def fn(data):
        out = net_forward(kind, weights, torch.from_numpy(np.ascontiguousarray(data)).float(), emulate_bf16=emulate_bf16)
        if kind == "hand":
            return out[0].numpy()
        return out[0][0].numpy(), out[1][0].numpy()

    return fn
