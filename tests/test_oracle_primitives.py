"""The numpy restatements of the third-party primitives against the libraries the reference calls."""
import numpy as np
import pytest

from oracle import openpose_oracle as O

cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("c", [19, 22, 26, 38, 52])
def test_float_cubic_resize_bit_exact_vs_cv2(c):
    rng = np.random.RandomState(c)
    src = rng.randn(23, 31, c).astype(np.float32)
    up = cv2.resize(src, (0, 0), fx=8, fy=8, interpolation=cv2.INTER_CUBIC)
    mine = O.resize_cubic(src, fx=8, fy=8)
    assert np.array_equal(up, mine)
    crop = np.ascontiguousarray(up[:181, :243])
    for dsize in [(320, 240), (333, 217), (97, 401)]:  # (W, H); 333*c is not a multiple of 4 for odd c
        assert np.array_equal(cv2.resize(crop, dsize, interpolation=cv2.INTER_CUBIC), O.resize_cubic(crop, dsize=dsize))


@pytest.mark.parametrize("hw", [(480, 640), (310, 458), (109, 109), (57, 91)])
@pytest.mark.parametrize("s", [0.5, 1.0, 1.5, 2.0])
def test_u8_cubic_resize_generic_path(hw, s):
    """Bit-exact against OpenCV's own generic code (IPP off); within one grey level of the IPP dispatch."""
    h, w = hw
    src = np.random.RandomState(h + w).randint(0, 256, (h, w, 3)).astype(np.uint8)
    scale = s * 368 / h
    mine = O.resize_cubic(src, fx=scale, fy=scale)
    with_ipp = cv2.resize(src, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
    had = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        generic = cv2.resize(src, (0, 0), fx=scale, fy=scale, interpolation=cv2.INTER_CUBIC)
    finally:
        cv2.ipp.setUseIPP(had)
    assert mine.shape == generic.shape == with_ipp.shape
    assert np.array_equal(mine, generic)
    d = np.abs(mine.astype(int) - with_ipp.astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 0.08


def test_gaussian_sigma3_bit_exact_vs_scipy():
    from scipy.ndimage import gaussian_filter

    rng = np.random.RandomState(0)
    for shape in [(64, 80), (7, 9), (240, 13), (25, 25)]:
        a = rng.rand(*shape)
        assert np.array_equal(gaussian_filter(a, sigma=3), O.gaussian_filter_sigma3(a))


def test_label8_vs_scipy():
    from scipy import ndimage

    rng = np.random.RandomState(1)
    for p in (0.2, 0.45, 0.7):
        b = (rng.rand(40, 57) < p).astype(np.uint8)
        ref, n = ndimage.label(b, structure=np.ones((3, 3), np.int32))
        lab, m = O.label8(b)
        assert n == m and np.array_equal(ref, lab)


def test_pad_and_npmax():
    img = np.arange(5 * 11 * 3, dtype=np.uint8).reshape(5, 11, 3)
    padded, pad = O.pad_right_down_corner(img, 8, 128)
    assert padded.shape == (8, 16, 3) and pad == [0, 0, 3, 5]
    assert np.array_equal(padded[:5, :11], img) and (padded[5:] == 128).all() and (padded[:, 11:] == 128).all()
    same, pad = O.pad_right_down_corner(np.zeros((16, 8, 3), np.uint8), 8, 128)
    assert same.shape == (16, 8, 3) and pad == [0, 0, 0, 0]
    a = np.array([[1.0, 5.0, 5.0], [5.0, 2.0, 0.0]])
    assert O.npmax(a) == (0, 1)
