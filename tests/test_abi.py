"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol include/islpose.h declares;
the product refuses to run without a GPU instead of falling back to anything."""
import ctypes
import os
import re

import pytest

import isl_b200
from isl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "islpose.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(islpose_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    names = _declared()
    assert len(names) >= 15
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), "libislpose.so does not export %s" % n
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"
    assert _lib.lib().islpose_abi_version() == _lib.ABI_VERSION == 3


def test_struct_sizes_match_the_header_layout():
    # LP64: pointers 8 bytes, int32 4 bytes, natural alignment - the same rules the C compiler applies
    assert ctypes.sizeof(_lib.Scale) == 24
    assert ctypes.sizeof(_lib.ConvDesc) == 208
    assert ctypes.sizeof(_lib.GroupBuffers) % 8 == 0
    assert ctypes.sizeof(_lib.HandCrop) == 8 + 4 * 24
    # and what the compiler made of include/islpose.h
    sizes = (ctypes.c_int32 * 4)()
    assert _lib.lib().islpose_struct_sizes(sizes) == 0
    assert list(sizes) == [ctypes.sizeof(_lib.Scale), ctypes.sizeof(_lib.ConvDesc), ctypes.sizeof(_lib.GroupBuffers), ctypes.sizeof(_lib.HandCrop)]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.IslposeError):
        isl_b200.Body({}, "coco")
    with pytest.raises(_lib.IslposeError):
        isl_b200.Hand({})
    from isl_b200 import translate as TR
    with pytest.raises(_lib.IslposeError):
        TR.Translator(TR.random_weights(7))


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.islpose_resize_pad_normalize(None, 1, 8, 8, 1.0, 8, 8, 8, 8, None, None, None) != 0
    assert b"null" in L.islpose_last_error()
    assert L.islpose_plan_add_conv(None, None) != 0
    # the classifier's entry point validates before it launches anything
    assert L.islpose_translate(None, 1, 20, 156, None, 0, 167, None, None) != 0 and b"bad argument" in L.islpose_last_error()
    assert L.islpose_translate(None, 0, 20, 156, None, 0, 167, None, None) == 0          # nothing to do
    assert L.islpose_translate_weight_floats(167) == 4 * 156 + 2 * (156 * 128 + 32 * 128 + 128) + 2 * (64 * 128 + 32 * 128 + 128) \
        + 64 * 32 + 4 * 32 + 32 * 32 + 4 * 32 + 32 * 167 + 167


def test_host_helpers_match_oracle():
    import numpy as np
    from oracle import openpose_oracle as O
    from isl_b200 import util
    assert np.array_equal(util.gaussian_weights(), O.gaussian_weights_sigma3())
    img = np.arange(5 * 11 * 3, dtype=np.uint8).reshape(5, 11, 3)
    a, pa = util.padRightDownCorner(img, 8, 128)
    b, pb = O.pad_right_down_corner(img, 8, 128)
    assert np.array_equal(a, b) and pa == pb
    g = np.load(os.path.join(ROOT, "tests", "golden", "body_body25_p12_s1.npz"))
    boxes = util.handDetect(g["candidate"], g["subset"], np.zeros((int(g["h"]), int(g["w"]), 3), np.uint8))
    assert np.array_equal(np.array([[b[0], b[1], b[2], int(b[3])] for b in boxes]), g["boxes"])
