"""Import the UNMODIFIED reference hot path (/root/reference/src/{model,body,hand,util}.py) in this container.

TEST INFRASTRUCTURE ONLY. This module exists to pin `oracle/openpose_oracle.py` against the reference itself
and to generate the committed fixtures under tests/golden/ (see tests/golden/make_golden.py). It needs
/root/reference, which does not exist on the GPU box, so nothing that runs there may import it.

Three shims, none of which edits the reference:
  1. matplotlib is imported at module top by body.py:6-7, hand.py:7-8, util.py:4-8 but never used on the hot
     path and is not installed -> empty stub modules are pre-seeded into sys.modules.
  2. skimage.measure.label (hand.py:10,67) is not installed -> provided by scipy.ndimage.label with a full
     3x3 structuring element (8-connectivity == connectivity=2 for 2-D input), same raster-order numbering.
  3. the trained weights are not in the tree -> objects are built with __new__ and a seeded random-init module.
"""
import os
import sys
import types

import numpy as np

# /root/reference in the build container; on the GPU box the copy of the four hot-path files that
# __graft_entry__.build() places under oracle/_ref/ (git-ignored, travels with the gpurun snapshot like built .so files)
_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("ISL_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")]
REFERENCE_ROOT = next((c for c in _CANDIDATES if c and os.path.isfile(os.path.join(c, "src", "body.py"))), "/root/reference")
HOT_PATH_FILES = ("model.py", "body.py", "hand.py", "util.py")


def vendor(dst=None):
    """Copies the reference's four hot-path files, unmodified, from /root/reference into oracle/_ref/src/ so that
    bench.py's reference arm can run the UNMODIFIED reference on the GPU box (where /root/reference does not exist).
    oracle/_ref/ is git-ignored: the files never enter this repository's history."""
    import shutil

    src_root = "/root/reference"
    if not os.path.isfile(os.path.join(src_root, "src", "body.py")):
        return False
    dst = dst or os.path.join(_HERE, "_ref")
    os.makedirs(os.path.join(dst, "src"), exist_ok=True)
    for f in HOT_PATH_FILES:
        shutil.copyfile(os.path.join(src_root, "src", f), os.path.join(dst, "src", f))
    open(os.path.join(dst, "src", "__init__.py"), "a").close()
    return True


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "body.py"))


def _stub_modules():
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.__path__ = []
        for sub in ("pyplot", "figure", "backends", "backends.backend_agg", "colors"):
            m = types.ModuleType("matplotlib." + sub)
            m.__path__ = []
            sys.modules["matplotlib." + sub] = m
        sys.modules["matplotlib.figure"].Figure = object
        sys.modules["matplotlib.backends.backend_agg"].FigureCanvasAgg = object
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
        mpl.colors = sys.modules["matplotlib.colors"]
        sys.modules["matplotlib"] = mpl
    if "skimage" not in sys.modules:
        from scipy import ndimage

        def label(binary, return_num=False, connectivity=None):
            lab, num = ndimage.label(binary, structure=np.ones((3, 3), dtype=np.int32))
            return (lab, num) if return_num else lab

        sk = types.ModuleType("skimage")
        sk.__path__ = []
        meas = types.ModuleType("skimage.measure")
        meas.label = label
        sk.measure = meas
        sys.modules["skimage"] = sk
        sys.modules["skimage.measure"] = meas


def load():
    """Returns the reference modules (model, body, hand, util)."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _stub_modules()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from src import body, hand, model, util  # noqa: E402  (the reference's own package)
    return model, body, hand, util


def make_body(model_type, net, scale_search=None):
    """Reference Body with a caller-supplied nn.Module (body.py:16-37 minus torch.load).

    body.py:40-41 hard-codes scale_search=[0.5] inside __call__; BASELINE's 4-scale configuration is the
    commented-out list, so the only way to run the unmodified __call__ with another list is to execute its
    source with that one literal replaced. That is done on a text copy in memory; the file is not touched.
    """
    _, body, _, _ = load()
    b = body.Body.__new__(body.Body)
    b.model = net
    b.model_type = model_type
    b.njoint, b.npaf = (26, 52) if model_type == "body25" else (19, 38)
    if scale_search is not None and list(scale_search) != [0.5]:
        import inspect
        import textwrap

        src = textwrap.dedent(inspect.getsource(body.Body.__call__))
        needle = "scale_search = [0.5]"
        assert src.count(needle) == 1
        src = src.replace(needle, "scale_search = %r" % (list(scale_search),))
        ns = {}
        exec(compile(src, "<reference body.py __call__ with scale_search patched>", "exec"), vars(body), ns)
        b.__class__ = type("BodyScales", (body.Body,), {"__call__": ns["__call__"]})
    return b


def make_hand(net):
    _, _, hand, _ = load()
    h = hand.Hand.__new__(hand.Hand)
    h.model = net
    return h


def reference_module(kind, flat):
    """The reference nn.Module with flat Caffe-named weights loaded exactly like Body/Hand.__init__ do
    (util.transfer + load_state_dict, body.py:35-36)."""
    rmodel, _, _, rutil = load()
    net = {"coco": rmodel.bodypose_model, "body25": rmodel.bodypose_25_model, "hand": rmodel.handpose_model}[kind]()
    net.load_state_dict(rutil.transfer(net, flat))
    return net.eval()
