"""Weight ingestion without torch (isl_b200/weights.py): the three on-disk formats of the reference's weight files."""
import numpy as np
import torch

import isl_b200  # noqa: F401
from isl_b200 import weights
from oracle import openpose_oracle as O
from packref import write_caffemodel


def _same(got, flat):
    assert sorted(got) == sorted(flat)
    for k, v in flat.items():
        assert got[k].dtype == np.float32 and got[k].shape == tuple(v.shape), k
        assert np.array_equal(got[k], v.numpy()), k


def test_torch_archives_are_read_without_torch(tmp_path):
    """torch.load(model_path) of body.py:35 / hand.py:20, both layouts torch has written over the years."""
    flat = O.make_flat_weights("hand", seed=1)
    for legacy in (False, True):
        p = str(tmp_path / ("w%d.pth" % legacy))
        torch.save(flat, p, _use_new_zipfile_serialization=not legacy)
        _same(weights.load_flat(p), flat)
    # nn.Parameter values, an OrderedDict, a {'state_dict': ...} wrapper, non-contiguous and double tensors
    import collections
    od = collections.OrderedDict((k, torch.nn.Parameter(v)) for k, v in flat.items())
    p = str(tmp_path / "od.pth")
    torch.save({"state_dict": od, "epoch": 3}, p)
    _same(weights.load_flat(p), flat)
    odd = {"a.weight": torch.arange(24, dtype=torch.float64).reshape(2, 3, 4).permute(2, 0, 1), "a.bias": torch.ones(3)[::2]}
    p = str(tmp_path / "odd.pth")
    torch.save(odd, p)
    got = weights.load_flat(p)
    assert np.array_equal(got["a.weight"], odd["a.weight"].float().numpy()) and np.array_equal(got["a.bias"], [1, 1])


def test_unpickler_refuses_code(tmp_path):
    import pickle

    import pytest

    class Evil(object):
        def __reduce__(self):
            return (print, ("executed",))

    p = str(tmp_path / "evil.pth")
    with open(p, "wb") as f:
        pickle.dump(0x1950A86A20F9469CFC6C, f, protocol=2)
        pickle.dump(1001, f, protocol=2)
        pickle.dump({}, f, protocol=2)
        pickle.dump({"x": Evil()}, f, protocol=2)
    with pytest.raises(weights.WeightFileError):
        weights.load_flat(p)


def test_caffemodel_protobuf(tmp_path):
    """caffemodel2pytorch.py:143-153: blobs[0] -> weight, blobs[1] -> bias, shapes from BlobShape or the legacy fields,
    `layer` and V1 `layers` messages alike."""
    flat = O.make_flat_weights("body25", seed=2)
    p = str(tmp_path / "pose_iter_584000.caffemodel")
    write_caffemodel(p, flat)
    got = weights.load_flat(p)
    assert sorted(got) == sorted(flat)
    for k, v in flat.items():
        assert np.array_equal(got[k].reshape(v.shape), v.numpy()), k
        if k.endswith(".weight") and v.ndim == 4:
            assert got[k].shape == tuple(v.shape)


def test_packed_blob_holds_the_bits_the_kernels_consume(tmp_path):
    flat = O.make_flat_weights("coco", seed=3)
    p = str(tmp_path / "body.islpose")
    weights.write_packed(p, flat)
    got = weights.load_flat(p)
    for k, v in flat.items():
        want = v.to(torch.bfloat16).float().numpy() if v.ndim == 4 else v.numpy()
        assert np.array_equal(got[k], want), k
    # rounding helper == torch's round-to-nearest-even, including ties and specials
    x = torch.tensor([1.0, 1.00390625, 1.01171875, -3.3e38, 3.4e38, 1e-40, float("inf"), 0.0, -0.0])
    bits = weights.f32_to_bf16_bits(x.numpy())
    assert np.array_equal(bits.view(np.int16), x.to(torch.bfloat16).view(torch.int16).numpy())
