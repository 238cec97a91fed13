"""B200-native OpenPose keypoint extraction (body + hand) behind the reference's call API.

    from isl_b200 import Body, Hand, util
    candidate, subset = Body(model_path, 'body25')(frame)                  # src/body.py:16,39
    for x, y, w, is_left in util.handDetect(candidate, subset, frame):     # src/util.py:242
        peaks = Hand(model_path)(frame[y:y+w, x:x+w, :])                   # src/hand.py:16,24

Everything under the three call signatures runs as hand-written sm_100a CUDA reached through the C ABI in
include/islpose.h (libislpose.so). There is no CPU fallback: without the built library or without a CUDA
device the constructors raise.
"""
from . import synth, tables, util, weights  # noqa: F401
from ._lib import IslposeError, configure  # noqa: F401
from .body import Body  # noqa: F401
from .hand import Hand  # noqa: F401
from .nets import PoseNet  # noqa: F401
from .translate import RollingTranslator, Translator  # noqa: F401
