"""Test helpers: a CPU restatement of the device weight packing (csrc/pack.cu) used as its checker, and a writer for
minimal Caffe NetParameter files (the reader under test is isl_b200.weights.read_caffemodel)."""
import torch


def up64(c):
    return (c + 63) // 64 * 64


def pack_reference(w, chan_map, in_c, first):
    """nn.Conv2d weight [cout, cin, k, k] float32 -> bf16 [k*k, cout, w_cin] in the buffer's channel order, zero beyond
    in_c up to a multiple of 64 (conv1_1: [1, cout, 32] with K index (ky*3+kx)*3+c)."""
    cout, cin, k, _ = w.shape
    if first:
        packed = torch.zeros((1, cout, in_c), dtype=torch.float32)
        packed[0, :, :27] = w.permute(0, 2, 3, 1).reshape(cout, 27)
        return packed.to(torch.bfloat16).contiguous()
    taps = w.permute(2, 3, 0, 1).reshape(k * k, cout, cin)
    w_cin = in_c if in_c % 64 == 0 else up64(in_c)
    packed = torch.zeros((k * k, cout, w_cin), dtype=torch.float32)
    if chan_map is None:
        assert cin == in_c
        packed[:, :, :cin] = taps
    else:
        idx = [(i, c) for i, c in enumerate(chan_map) if c is not None]
        dst = torch.tensor([i for i, _ in idx], dtype=torch.long)
        src = torch.tensor([c for _, c in idx], dtype=torch.long)
        assert sorted(src.tolist()) == list(range(cin))
        packed[:, :, dst] = taps[:, :, src]
    return packed.to(torch.bfloat16).contiguous()


def _varint(v):
    out = b""
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out += bytes([b | 0x80])
        else:
            return out + bytes([b])


def _ld(field, payload):
    return _varint(field << 3 | 2) + _varint(len(payload)) + payload


def write_caffemodel(path, flat, v1_every=5):
    """A minimal Caffe NetParameter: one `layer` (or, for every v1_every-th, a V1 `layers`) message per Caffe layer
    with its blobs - weight [cout,cin,k,k], bias [cout] - shapes in BlobShape or the legacy num/channels/height/width."""
    names = []
    for k in flat:
        n = k.rsplit(".", 1)[0]
        if n not in names:
            names.append(n)
    out = b""
    for i, n in enumerate(names):
        blobs = b""
        for suffix in ("weight", "bias"):
            t = flat.get("%s.%s" % (n, suffix))
            if t is None:
                continue
            a = t.numpy().astype("<f4")
            if i % 2 == 0:
                shape = _ld(7, _ld(1, b"".join(_varint(d) for d in a.shape)))
            else:
                dims = ([1] * (4 - a.ndim) + list(a.shape))
                shape = b"".join(_varint(f << 3) + _varint(d) for f, d in zip((1, 2, 3, 4), dims))
            blobs_one = shape + _ld(5, a.tobytes())
            blobs += _ld(7 if i % v1_every else 6, blobs_one)
        if i % v1_every:
            out += _ld(100, _ld(1, n.encode()) + _ld(2, b"Convolution") + blobs)
        else:
            out += _ld(2, _ld(4, n.encode()) + blobs)
    open(path, "wb").write(_ld(1, b"synthetic") + out)
