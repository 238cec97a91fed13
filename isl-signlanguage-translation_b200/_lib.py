"""ctypes binding of libislpose.so (C ABI declared in include/islpose.h).

There is no CPU fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes as C
import os
import sys
import warnings


def configure(max_connections=32):
    """Opt-in, process-wide: asks the CUDA driver for `max_connections` hardware work queues
    (CUDA_DEVICE_MAX_CONNECTIONS). Scales, pipeline lanes and hand crops run on a few dozen streams; with the default
    of 8 queues unrelated streams share a queue and serialise, which costs throughput in KeypointExtractor.pipeline().
    The variable is only read when the CUDA context is created, so call this before the first CUDA call of the
    process (bench.py and the tools do). Returns False, with a warning, when a context already exists or the
    variable is already set to something else; nothing is changed in that case."""
    cur = os.environ.get("CUDA_DEVICE_MAX_CONNECTIONS")
    if cur is not None:
        if cur != str(max_connections):
            warnings.warn("CUDA_DEVICE_MAX_CONNECTIONS is already %s; leaving it" % cur)
            return False
        return True
    torch = sys.modules.get("torch")
    if torch is not None and torch.cuda.is_initialized():
        warnings.warn("isl_b200.configure() called after the CUDA context was created: CUDA_DEVICE_MAX_CONNECTIONS would "
                      "have no effect, stream overlap in KeypointExtractor.pipeline() is limited to 8 hardware queues")
        return False
    os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = str(max_connections)
    return True


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libislpose.so")

MAX_SCALES = 8
ABI_VERSION = 3


class ConvDesc(C.Structure):
    _fields_ = [("in_", C.c_void_p), ("in_c", C.c_int32), ("in_cstride", C.c_int32), ("in_c_readable", C.c_int32),
                ("w_cin", C.c_int32), ("n", C.c_int32), ("h", C.c_int32),
                ("w", C.c_int32), ("weights", C.c_void_p), ("cout", C.c_int32), ("ksize", C.c_int32),
                ("bias", C.c_void_p), ("slope", C.c_void_p), ("out_bf16", C.c_void_p), ("out_cstride", C.c_int32),
                ("out_f32", C.c_void_p), ("out_f32_channels", C.c_int32), ("n_tile", C.c_int32), ("stages", C.c_int32),
                ("tile_w", C.c_int32), ("tile_h", C.c_int32), ("pool", C.c_int32), ("sm_budget", C.c_int32),
                ("weights2", C.c_void_p), ("cout2", C.c_int32), ("bias2", C.c_void_p), ("slope2", C.c_void_p),
                ("out2_bf16", C.c_void_p), ("out2_cstride", C.c_int32), ("out2b_bf16", C.c_void_p), ("out2b_cstride", C.c_int32),
                ("out2_f32", C.c_void_p), ("out2_f32_channels", C.c_int32)]


class Scale(C.Structure):
    _fields_ = [("lowres", C.c_void_p), ("gh", C.c_int32), ("gw", C.c_int32), ("hc", C.c_int32), ("wc", C.c_int32)]


class HandCrop(C.Structure):
    _fields_ = [("h", C.c_int32), ("w", C.c_int32), ("scales", Scale * 4)]


OVERFLOW_PEAKS, OVERFLOW_CANDIDATES, OVERFLOW_PAIRS, OVERFLOW_PERSONS = 1, 2, 4, 8


class GroupBuffers(C.Structure):
    _fields_ = [("cap", C.c_int32), ("counts", C.c_void_p), ("keys", C.c_void_p), ("scores", C.c_void_p),
                ("pair_cap", C.c_int64), ("pair_score", C.c_void_p), ("end_paf", C.c_void_p), ("conn_count", C.c_void_p), ("conn_ij", C.c_void_p),
                ("conn_score", C.c_void_p), ("owner", C.c_void_p), ("max_cand", C.c_int32), ("candidate", C.c_void_p), ("n_cand", C.c_void_p),
                ("max_person", C.c_int32), ("subset", C.c_void_p), ("n_person", C.c_void_p), ("overflow", C.c_void_p)]


# every symbol include/islpose.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "islpose_abi_version": (C.c_int, []),
    "islpose_last_error": (C.c_char_p, []),
    "islpose_launch_count": (C.c_int64, []),
    "islpose_struct_sizes": (C.c_int, [C.POINTER(C.c_int32)]),
    "islpose_pack_conv_weights": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                            C.c_void_p, C.c_void_p]),
    "islpose_plan_create": (C.c_int, [C.POINTER(C.c_void_p)]),
    "islpose_plan_destroy": (C.c_int, [C.c_void_p]),
    "islpose_plan_add_conv": (C.c_int, [C.c_void_p, C.POINTER(ConvDesc)]),
    "islpose_plan_add_first_conv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                            C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "islpose_plan_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "islpose_plan_set_graph": (C.c_int, [C.c_void_p, C.c_int32]),
    "islpose_plan_graph_state": (C.c_int32, [C.c_void_p]),
    "islpose_plan_graph_note": (C.c_char_p, [C.c_void_p]),
    "islpose_plan_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "islpose_plan_num_launches": (C.c_int32, [C.c_void_p]),
    "islpose_plan_conv_flops": (C.c_double, [C.c_void_p]),
    "islpose_resize_pad_normalize": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32,
                                               C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "islpose_maps_workspace_floats": (C.c_int64, [C.POINTER(Scale), C.c_int32, C.c_int32, C.c_int32]),
    "islpose_maps_accumulate": (C.c_int, [C.POINTER(Scale), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "islpose_body_peaks": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_double, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "islpose_body_group": (C.c_int, [C.POINTER(Scale), C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double,
                                     C.c_int32, C.POINTER(GroupBuffers), C.c_void_p]),
    "islpose_body_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p]),
    "islpose_hand_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "islpose_translate_weight_floats": (C.c_int64, [C.c_int32]),
    "islpose_translate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p,
                                    C.c_void_p]),
    "islpose_hand_workspace_bytes": (C.c_int64, [C.POINTER(HandCrop), C.c_int32]),
    "islpose_hand_keypoints": (C.c_int, [C.POINTER(HandCrop), C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_double,
                                         C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "islpose_hand_peaks": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.c_double,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib = None


class IslposeError(RuntimeError):
    pass


def lib():
    """The loaded library. Raises IslposeError when libislpose.so has not been built (no fallback exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise IslposeError("libislpose.so is missing at %s - build it with `python -c 'import __graft_entry__ as g; "
                               "g.build()'` (nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.islpose_abi_version() != ABI_VERSION:
            raise IslposeError("libislpose.so has ABI version %d, expected %d (rebuild it)" % (handle.islpose_abi_version(),
                                                                                            ABI_VERSION))
        sizes = (C.c_int32 * 4)()
        handle.islpose_struct_sizes(sizes)
        mine = [C.sizeof(Scale), C.sizeof(ConvDesc), C.sizeof(GroupBuffers), C.sizeof(HandCrop)]
        if list(sizes) != mine:
            raise IslposeError("struct layouts of libislpose.so %s and of this binding %s differ (rebuild the library)" % (list(sizes), mine))
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        raise IslposeError("%s failed: %s" % (what, lib().islpose_last_error().decode("utf-8", "replace")))


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)
