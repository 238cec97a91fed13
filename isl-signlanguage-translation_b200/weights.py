"""Weight ingestion without torch: the reference's weight files -> flat dict {caffe_name.weight|bias: float32 array}.

The reference loads `torch.load(model_path)` and renames the keys with util.transfer (src/body.py:35-36, src/hand.py:20,
src/util.py:35-44); its files are flat dicts of Caffe layer names (`./model/body_pose_model.pth`,
`./model/pose_iter_584000.caffemodel.pt`, `model/hand_pose_model.pth`, demo.py:12-17) produced from `.caffemodel`
protobufs by caffemodel2pytorch (caffemodel2pytorch/caffemodel2pytorch.py:61-160: blobs[0] -> weight, blobs[1] -> bias).
Here the three on-disk formats are parsed directly:

  read_pth         torch.save archives, both the zip layout (torch >= 1.6) and the legacy stream layout, with a
                   restricted unpickler that only knows tensors, storages and ordered dicts - no torch import, no
                   arbitrary code execution
  read_caffemodel  Caffe NetParameter protobuf wire format: `layer` (field 100) / V1 `layers` (field 2) -> name, blobs
  read_packed      this package's own compact blob (write_packed): bf16 weights + float32 bias / PReLU slopes, i.e.
                   exactly the bits the kernels consume, half the size of the float32 file

load_flat() picks by content. The arrays go to the device as float32 parameters of nets.PoseNet and are packed into the
kernels' operand layout there (csrc/pack.cu).
"""
import io
import json
import os
import pickle
import struct
import zipfile

import numpy as np

_DTYPES = {"FloatStorage": np.float32, "DoubleStorage": np.float64, "HalfStorage": np.float16, "LongStorage": np.int64,
           "IntStorage": np.int32, "ShortStorage": np.int16, "CharStorage": np.int8, "ByteStorage": np.uint8,
           "BoolStorage": np.bool_, "BFloat16Storage": "bfloat16"}
PACKED_MAGIC = b"ISLPOSEW1\n"


class WeightFileError(ValueError):
    pass


def _bf16_to_f32(raw_u16):
    return (raw_u16.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16_bits(a):
    """float32 array -> uint16 bf16 bit patterns, round-to-nearest-even (what torch's .to(bfloat16) and the device
    packing kernel do); NaN stays NaN."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint16)
    nan = np.isnan(a)
    if nan.any():
        rounded = np.where(nan, np.uint16(0x7FC0), rounded)
    return rounded


class _Storage(object):
    def __init__(self, dtype, key, numel):
        self.dtype, self.key, self.numel, self.data = dtype, key, numel, None

    def array(self):
        if self.data is None:
            raise WeightFileError("storage %r was never filled" % (self.key,))
        if self.dtype == "bfloat16":
            return _bf16_to_f32(np.frombuffer(self.data, dtype=np.uint16))
        return np.frombuffer(self.data, dtype=self.dtype)


class _StorageType(object):
    def __init__(self, name):
        self.name = name


def _rebuild_tensor_v2(storage, storage_offset, size, stride, requires_grad=False, backward_hooks=None, metadata=None):
    size, stride = tuple(size), tuple(stride)
    if storage.data is None:   # first pass over a legacy stream: the storages follow the object pickle
        return np.zeros(size, dtype=np.float32)
    flat = storage.array()
    if len(size) == 0:
        return np.array(flat[storage_offset])
    view = np.lib.stride_tricks.as_strided(flat[storage_offset:], shape=size, strides=tuple(s * flat.itemsize for s in stride))
    return np.array(view)   # an owned, contiguous copy


def _rebuild_tensor(storage, storage_offset, size, stride):
    return _rebuild_tensor_v2(storage, storage_offset, size, stride)


def _rebuild_parameter(data, requires_grad=False, backward_hooks=None):
    return data


class _Unpickler(pickle.Unpickler):
    """Knows exactly what a saved state dict of tensors needs; anything else is refused."""

    def __init__(self, f, storages, fill=None):
        super().__init__(f, encoding="utf-8")
        self.storages = storages
        self.fill = fill   # zip archives: reads a storage's bytes as soon as it is referenced

    def find_class(self, module, name):
        if module == "collections" and name == "OrderedDict":
            import collections
            return collections.OrderedDict
        if module == "torch._utils" and name == "_rebuild_tensor_v2":
            return _rebuild_tensor_v2
        if module == "torch._utils" and name == "_rebuild_tensor":
            return _rebuild_tensor
        if module == "torch._utils" and name == "_rebuild_parameter":
            return _rebuild_parameter
        if module == "torch" and name in _DTYPES:
            return _StorageType(name)
        if module == "torch.serialization" and name == "_get_layout":
            return lambda *a: None
        raise WeightFileError("weight file references %s.%s, which is not part of a flat tensor dict" % (module, name))

    def persistent_load(self, pid):
        if not isinstance(pid, tuple) or pid[0] != "storage":
            raise WeightFileError("unknown persistent id %r" % (pid,))
        stype, key, numel = pid[1], str(pid[2]), pid[4]
        name = stype.name if isinstance(stype, _StorageType) else getattr(stype, "__name__", str(stype))
        if name not in _DTYPES:
            raise WeightFileError("unsupported storage type %s" % name)
        st = self.storages.get(key)
        if st is None:
            st = self.storages[key] = _Storage(_DTYPES[name], key, numel)
        if self.fill is not None and st.data is None:
            st.data = self.fill(st.key)
        return st


def _as_flat_dict(obj):
    if not hasattr(obj, "items"):
        raise WeightFileError("weight file holds a %s, expected a dict of tensors" % type(obj).__name__)
    if "state_dict" in obj and hasattr(obj["state_dict"], "items"):
        obj = obj["state_dict"]
    out = {}
    for k, v in obj.items():
        if isinstance(v, np.ndarray):
            out[str(k)] = v.astype(np.float32, copy=False)
    if not out:
        raise WeightFileError("weight file holds no tensors")
    return out


def read_pth(path):
    """torch.save()'d flat dict -> {name: float32 ndarray}, without importing torch."""
    storages = {}
    if zipfile.is_zipfile(path):
        with zipfile.ZipFile(path) as z:
            names = z.namelist()
            pkl = [n for n in names if n.endswith("data.pkl")]
            if not pkl:
                raise WeightFileError("%s: zip archive without data.pkl" % path)
            root = pkl[0][:-len("data.pkl")]
            up = _Unpickler(io.BytesIO(z.read(pkl[0])), storages, fill=lambda key: z.read(root + "data/" + key))
            return _as_flat_dict(up.load())
    with open(path, "rb") as f:
        magic = pickle.load(f)
        if magic != 0x1950A86A20F9469CFC6C:
            raise WeightFileError("%s is neither a torch zip archive nor a legacy torch.save stream" % path)
        pickle.load(f)   # protocol version
        pickle.load(f)   # sys info
        start = f.tell()
        # the tensors reference storages that are only filled after the object pickle: two passes
        up = _Unpickler(f, storages)
        up.load()
        keys = pickle.load(f)
        for key in keys:
            st = storages.get(str(key))
            numel, = struct.unpack("<q", f.read(8))
            itemsize = 2 if st is None or st.dtype == "bfloat16" else np.dtype(st.dtype).itemsize
            data = f.read(numel * itemsize)
            if st is not None:
                st.data = data
        f.seek(start)
        return _as_flat_dict(_Unpickler(f, storages).load())


# ---- Caffe protobuf wire format ---------------------------------------------------------------------------
def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """Yields (field number, wire type, value) of one protobuf message; length-delimited values as memoryviews."""
    pos, end = 0, len(buf)
    while pos < end:
        tag, pos = _varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            val, pos = buf[pos:pos + 4], pos + 4
        else:
            raise WeightFileError("unsupported protobuf wire type %d" % wt)
        yield num, wt, val


def _blob(buf):
    """BlobProto: shape = 7 {dim = 1}, data = 5 (packed float), double_data = 8, legacy num/channels/height/width = 1..4."""
    dims, legacy, chunks, dchunks = [], {}, [], []
    for num, wt, val in _fields(buf):
        if num == 7 and wt == 2:
            for n2, w2, v2 in _fields(val):
                if n2 == 1 and w2 == 2:
                    p = 0
                    while p < len(v2):
                        d, p = _varint(v2, p)
                        dims.append(d)
                elif n2 == 1:
                    dims.append(v2)
        elif num == 5:
            chunks.append(np.frombuffer(bytes(val), dtype="<f4"))
        elif num == 8:
            dchunks.append(np.frombuffer(bytes(val), dtype="<f8"))
        elif num in (1, 2, 3, 4) and wt == 0:
            legacy[num] = val
    data = np.concatenate(chunks) if chunks else (np.concatenate(dchunks).astype(np.float32) if dchunks else np.zeros(0, np.float32))
    if not dims and legacy:
        dims = [legacy.get(i, 1) for i in (1, 2, 3, 4)]
    return data.reshape(dims) if dims else data


def read_caffemodel(path):
    """Caffe NetParameter -> {layer.weight: blobs[0], layer.bias: blobs[1]} (caffemodel2pytorch.py:149-153)."""
    buf = memoryview(open(path, "rb").read())
    out = {}
    for num, wt, val in _fields(buf):
        if wt != 2 or num not in (100, 2):
            continue
        name_field, blob_field = (1, 7) if num == 100 else (4, 6)
        name, blobs = None, []
        for n2, w2, v2 in _fields(val):
            if n2 == name_field and w2 == 2:
                name = bytes(v2).decode("utf-8")
            elif n2 == blob_field and w2 == 2:
                blobs.append(_blob(v2))
        if name is None:
            continue
        for suffix, b in zip(("weight", "bias"), blobs):
            out["%s.%s" % (name, suffix)] = np.ascontiguousarray(b.reshape(-1) if suffix == "bias" else b, dtype=np.float32)
    if not out:
        raise WeightFileError("%s holds no Caffe layers with blobs" % path)
    return out


# ---- packed blob ------------------------------------------------------------------------------------------
def write_packed(path, flat):
    """Flat dict -> compact blob: conv weights as bf16 bit patterns (the rounding the kernels apply anyway), everything
    else float32. A PoseNet built from the blob computes bit-identical results to one built from the float32 file."""
    table, payload, off = [], [], 0
    for name, a in flat.items():
        a = np.ascontiguousarray(np.asarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a), dtype=np.float32)
        as_bf16 = a.ndim == 4
        raw = f32_to_bf16_bits(a).tobytes() if as_bf16 else a.tobytes()
        table.append({"name": name, "shape": list(a.shape), "dtype": "bf16" if as_bf16 else "f32", "offset": off, "bytes": len(raw)})
        payload.append(raw)
        off += len(raw) + (-len(raw)) % 16
        payload.append(b"\0" * ((-len(raw)) % 16))
    header = json.dumps({"tensors": table}).encode("utf-8")
    with open(path, "wb") as f:
        f.write(PACKED_MAGIC)
        f.write(struct.pack("<q", len(header)))
        f.write(header)
        f.write(b"".join(payload))


def read_packed(path):
    with open(path, "rb") as f:
        if f.read(len(PACKED_MAGIC)) != PACKED_MAGIC:
            raise WeightFileError("%s is not a packed islpose weight blob" % path)
        hlen, = struct.unpack("<q", f.read(8))
        table = json.loads(f.read(hlen).decode("utf-8"))["tensors"]
        base = f.tell()
        body = np.fromfile(f, dtype=np.uint8)
    out = {}
    for t in table:
        raw = body[t["offset"]:t["offset"] + t["bytes"]]
        a = _bf16_to_f32(raw.view(np.uint16)) if t["dtype"] == "bf16" else raw.view(np.float32)
        out[t["name"]] = a.reshape(t["shape"]).copy()
    del base
    return out


def load_flat(model_path):
    """What Body / Hand accept as `model_path`: a dict that is already flat, or a file in one of the three formats."""
    if isinstance(model_path, dict):
        return model_path
    path = os.fspath(model_path)
    with open(path, "rb") as f:
        head = f.read(len(PACKED_MAGIC))
    if head == PACKED_MAGIC:
        return read_packed(path)
    if zipfile.is_zipfile(path) or head[:2] == b"\x80\x02":
        return read_pth(path)
    return read_caffemodel(path)
