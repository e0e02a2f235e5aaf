"""ctypes binding of libgnnmp.so (the C ABI declared in include/gnnmp.h).

There is no CPU fallback: if the shared library is missing, or no sm_100 device is visible when a
compute entry point is needed, this module raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GNNMP_LIB_PATH") or os.path.join(_HERE, "libgnnmp.so")   # (override: A/B builds of the same library)

GMP_DTYPE_F32 = 0
GMP_DTYPE_F64 = 1

# name -> (restype, argtypes); mirrors include/gnnmp.h one to one (checked by tests/test_abi.py)
SIGNATURES = {
    "gmp_last_error": (c_char_p, []),
    "gmp_version": (c_char_p, []),
    "gmp_device_ok": (c_int, [c_int]),
    "gmp_create": (c_void_p, [c_int]),
    "gmp_destroy": (None, [c_void_p]),
    "gmp_explorer_init": (c_int, [c_void_p, c_int, c_int, c_int]),
    "gmp_explorer_set_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "gmp_explorer_finalize": (c_int, [c_void_p]),
    "gmp_explorer_set_edge_feature_mode": (c_int, [c_void_p, c_int]),
    "gmp_explorer_workspace_bytes": (c_int64, [c_void_p, c_int64, c_int64, c_int64, c_int64]),
    "gmp_explorer_forward": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "gmp_explorer_bad_edges": (c_int, [c_void_p, c_void_p]),
    "gmp_set_timing": (c_int, [c_void_p, c_int]),
    "gmp_get_timings": (c_int, [c_void_p, c_void_p, c_int]),
    "gmp_knn_graph_max_edges": (c_int64, [c_int64, c_int]),
    "gmp_knn_graph_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64, c_int]),
    "gmp_knn_graph": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                              c_void_p, c_void_p, c_int64, c_void_p]),
    "gmp_maze_state_fp": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gmp_maze_edge_fp": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gmp_maze_edge_fp_graph": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                       c_void_p, c_void_p, c_void_p, c_void_p]),
    "gmp_maze3_state_fp": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gmp_maze3_edge_fp": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gmp_tree_search_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64]),
    "gmp_maze_tree_search": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_int64, c_void_p]),
    "gmp_arm_tree_search": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, ctypes.c_double, c_void_p, c_int64, c_int64, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p,
                                    c_int64, c_void_p]),
    "gmp_search_result_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "gmp_maze_sample_points": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, ctypes.c_uint64, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p]),
    "gmp_maze_steer_rounds": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_double, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "gmp_arm_model_count": (c_int, []),
    "gmp_arm_model_info": (c_int, [c_int, c_void_p, c_void_p, c_void_p]),
    "gmp_arm_state_fp": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gmp_arm_edge_fp": (c_int, [c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int64, ctypes.c_double,
                                c_void_p, c_void_p, c_void_p]),
    "gmp_arm_edge_fp_graph": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                      c_void_p, c_void_p, ctypes.c_double, c_void_p, c_void_p, c_void_p]),
    "gmp_arm_edge_fp_graph_cached": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                             c_void_p, c_void_p, ctypes.c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gmp_arm_edge_graph_workspace_bytes": (c_int64, [c_int64, c_int64]),
    "gmp_arm_edge_fp_graph_fast": (c_int, [c_int, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                           c_void_p, c_void_p, c_int, ctypes.c_double, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "gmp_smoother_init": (c_int, [c_void_p, c_int, c_int]),
    "gmp_smoother_set_tensor": (c_int, [c_void_p, c_char_p, c_void_p, c_int64]),
    "gmp_smoother_finalize": (c_int, [c_void_p]),
    "gmp_smoother_workspace_bytes": (c_int64, [c_void_p, c_int64, c_int64, c_int64, c_int64]),
    "gmp_smoother_forward": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_float, c_int, c_void_p, c_void_p, c_int64, c_void_p]),
    "gmp_result_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "gmp_edge_index_narrow": (c_int, [c_void_p, c_int64, c_int64, c_int, c_void_p, c_int64, c_void_p]),
    "gmp_post_to_host": (c_int, [c_void_p, c_int64, c_void_p, c_void_p]),
}

_lib = None


class GnnmpError(RuntimeError):
    pass


def load():
    """Load libgnnmp.so (built in-tree by ``__graft_entry__.build()`` / ``make -C csrc``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GnnmpError(
            "libgnnmp.so not found at %s -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise GnnmpError("libgnnmp error %d: %s" % (rc, load().gmp_last_error().decode()))


def ptr(t):
    """Device/host pointer of a torch tensor or numpy array (None -> NULL)."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return t.data_ptr()
    return t.ctypes.data


def stream_ptr(device):
    import torch
    return torch.cuda.current_stream(device).cuda_stream


_handles = {}


def handle(device_index):
    """One library handle per (process, device, owner) is created by the model classes; this returns a
    shared handle for stateless entry points that still want a device check."""
    if device_index not in _handles:
        lib = load()
        h = lib.gmp_create(device_index)
        if not h:
            raise GnnmpError(lib.gmp_last_error().decode())
        _handles[device_index] = h
    return _handles[device_index]


def require_cuda(t, what):
    if not (hasattr(t, "is_cuda") and t.is_cuda):
        raise GnnmpError("%s must be a CUDA tensor: the B200 path has no CPU fallback" % what)
