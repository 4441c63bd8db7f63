"""ctypes binding of libb200match.so (the C ABI declared in include/b200m.h).

The product path has no CPU fallback: if the shared library is missing or the CUDA
device is not an sm_100 part, every call raises.  ``build()`` compiles the library
in-tree with nvcc for sm_100a (it cross-compiles without a GPU).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200M_LIB: developer override (A/B timing of two builds); the in-tree library is the product
LIB_PATH = os.environ.get("B200M_LIB") or os.path.join(_HERE, "libb200match.so")
CSRC = os.path.join(_HERE, "csrc")

MAX_KENC = 8
MAX_GNN = 64

ERR_RAGGED = -5


class Config(C.Structure):
    _fields_ = [("descriptor_dim", C.c_int), ("nms_radius", C.c_int), ("keypoint_threshold", C.c_float),
                ("max_keypoints", C.c_int), ("remove_borders", C.c_int), ("align_corners", C.c_int),
                ("n_kenc", C.c_int), ("kenc", C.c_int * MAX_KENC), ("n_gnn_layers", C.c_int),
                ("gnn_cross", C.c_int * MAX_GNN), ("sinkhorn_iterations", C.c_int),
                ("match_threshold", C.c_float)]


# name -> (restype, argtypes); kept in one table so tests can check every symbol of b200m.h is exported
_P = C.c_void_p
_I = C.c_int
_Z = C.c_size_t
SIGNATURES = {
    "b200m_last_error": (C.c_char_p, []),
    "b200m_version": (_I, []),
    "b200m_create": (_I, [C.POINTER(Config), _I, C.POINTER(_P)]),
    "b200m_destroy": (None, [_P]),
    "b200m_set_tensor": (_I, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), _I]),
    "b200m_pack": (_I, [_P, _P]),
    "b200m_keypoint_capacity": (_I, [_P, _I, _I]),
    "b200m_superpoint_workspace_bytes": (_Z, [_P, _I, _I, _I]),
    "b200m_superpoint_forward": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "b200m_superpoint_forward_u8": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "b200m_superpoint_dense": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "b200m_detector_post": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _I, _P, _Z, _P]),
    "b200m_sample_descriptors": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "b200m_superglue_workspace_bytes": (_Z, [_P, _I, _I, _I]),
    "b200m_superglue_forward": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I,
                                     _P, _P, _P, _P, _P, _Z, _P]),
    "b200m_keypoint_encode": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "b200m_gnn": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _P, _P, _P, _Z, _P]),
    "b200m_score_matrix": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _Z, _P]),
    "b200m_sinkhorn": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _Z, _P]),
    "b200m_knn_ratio_match": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, C.c_float, _P, _P, _P, _P]),
    "b200m_estimate_affine_partial": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, C.c_double, _I, C.c_double, _I, _P, _P, _P,
                                           _P]),
    "b200m_warp_affine": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _I, _I, _P]),
    "b200m_match_select": (_I, [_P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _Z, _P]),
    "b200m_matching_workspace_bytes": (_Z, [_P, _I, _I, _I]),
    "b200m_matching_forward": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I,
                                    _P, _P, _P, _P, _P, _Z, _P]),
    "b200m_matching_forward_u8": (_I, [_P, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _I,
                                       _P, _P, _P, _P, _P, _Z, _P]),
    "b200m_resize_linear_u8": (_I, [_P, _P, _I, _I, _I, _P, _I, _I, _P]),
    "b200m_pack_match_wire": (_I, [_P, _P, _P, _I, _I, _I, _I, _P, _P]),
    "b200m_unpack_match_wire": (_I, [_P, _P, _I, _I, _I, _I, _P, _P, _P]),
    "b200m_debug_conv_layer": (_I, [_P, _I, _I, _P, _P, _I, _I, _I, _P]),
    "b200m_debug_attention": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "b200m_launch_count": (C.c_longlong, [_P]),
    "b200m_graph_replay_count": (C.c_longlong, [_P]),
    "b200m_profile_begin": (_I, [_P, _I]),
    "b200m_profile_end": (_I, [_P, C.c_char_p, _Z]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libb200match.so for sm_100a with the in-tree Makefile; returns its path."""
    r = subprocess.run(["make", "-C", CSRC, "-j", str(os.cpu_count() or 4)], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libb200match.so failed (nvcc / make)")
    return LIB_PATH


def load():
    """dlopen the library and set the prototypes.  Raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "or `make -C image_matching_b200/csrc` first (the CUDA path has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().b200m_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = "libb200match"):
    if rc == 0:
        return
    msg = last_error()
    if rc == -1:
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what} failed ({rc}): {msg}")
