"""ctypes binding of libgeossl_b200.so (C ABI in include/geossl_b200.h).

The shared library is plain nvcc output (no torch types in its signatures); Python passes
``tensor.data_ptr()`` and the current CUDA stream.  There is NO fallback: if the library is missing
or a call fails a RuntimeError is raised.
"""
import ctypes
import glob
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libgeossl_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--threads", "0"]

_lock = threading.Lock()
_lib = None

c_p = ctypes.c_void_p
c_i64 = ctypes.c_int64
c_int = ctypes.c_int
c_f = ctypes.c_float


class DdmPtrs(ctypes.Structure):
    """geossl_ddm_params / geossl_ddm_grads (ten device pointers, same order)."""
    _fields_ = [(n, c_p) for n in ("in_w0", "in_b0", "in_w1", "in_b1", "out_w0", "out_b0", "out_w1", "out_b1",
                                   "out_w2", "out_b2")]


class ChainStage(ctypes.Structure):
    """geossl_chain_stage (include/geossl_b200.h)."""
    _fields_ = [("weight_image", c_p), ("bias", c_p), ("act_grad_input", c_p), ("residual", c_p), ("store", c_p), ("act_next", c_int)]


class ChainStageEx(ctypes.Structure):
    """geossl_chain_stage_ex (include/geossl_b200.h)."""
    _fields_ = [("weight_image", c_p), ("bias", c_p), ("act_grad_input", c_p), ("ldz", c_i64), ("residual", c_p), ("ldr", c_i64),
                ("store", c_p), ("ld_store", c_i64), ("x", c_p), ("ldx", c_i64), ("act_next", c_int), ("x_act", c_int),
                ("keep", c_int), ("accumulate", c_int), ("partial", c_int)]


class WgradProblem(ctypes.Structure):
    """geossl_wgrad_problem (include/geossl_b200.h)."""
    _fields_ = [("grad_y", c_p), ("ld_dy", c_i64), ("x", c_p), ("ld_x", c_i64), ("workspace", c_p), ("grad_weight", c_p),
                ("ld_gw", c_int), ("grad_bias", c_p), ("pre_act", c_int), ("x_cols", c_int)]


ABI_VERSION = 4          # must equal GEOSSL_ABI_VERSION in include/geossl_b200.h

_SIGNATURES = {
    "geossl_abi_version": (c_int, []),
    "geossl_last_error": (ctypes.c_char_p, []),
    "geossl_launch_count": (c_i64, [c_int]),
    "geossl_graph_ptr": (c_int, [c_p, c_i64, c_i64, c_p, c_p]),
    "geossl_radius_csr": (c_int, [c_p, c_p, c_p, c_i64, c_f, c_int, c_i64, c_p, c_p, c_p, c_p, c_p, c_int, c_p, c_p, c_p, c_int, c_p]),
    "geossl_radius_cell_keys": (c_int, [c_p, c_p, c_p, c_i64, c_i64, c_f, c_p, c_p, c_p]),
    "geossl_csr_to_edge_index": (c_int, [c_p, c_p, c_i64, c_p, c_p]),
    "geossl_csr_transpose": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p]),
    "geossl_super_edges": (c_int, [c_p, c_p, c_i64, c_i64, c_int, c_i64, c_p, c_p, c_p]),
    "geossl_rowptr_from_sorted": (c_int, [c_p, c_i64, c_i64, c_p, c_p]),
    "geossl_filter_fwd": (c_int, [c_p, c_p, c_i64, c_p, c_f, c_f, c_int, c_int, c_p, c_p, c_p, c_p, c_p, c_p]),
    "geossl_filter_fwd_tc": (c_int, [c_p, c_p, c_i64, c_p, c_f, c_f, c_int, c_int, c_p, c_p, c_p, c_p, c_p, c_int, c_p]),
    "geossl_debug_set_trace": (c_int, [c_p]),
    "geossl_debug_set_trace_bwd": (c_int, [c_p]),
    "geossl_debug_set_trace_head": (c_int, [c_p]),
    "geossl_debug_set_trace_linear": (c_int, [c_p]),
    "geossl_tc_selftest": (c_int, [c_int, c_int, c_p, c_p, c_int, c_int, c_p, c_p]),
    "geossl_cfconv_fwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_i64, c_int, c_p, c_p]),
    "geossl_cfconv_bwd_x": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_int, c_p, c_p]),
    "geossl_ssp_family": (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p]),
    "geossl_row_scale": (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p]),
    "geossl_row_dot": (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p]),
    "geossl_cfconv_pair_product": (c_int, [c_p, c_p, c_p, c_i64, c_int, c_p, c_p]),
    "geossl_cfconv_pairs_max_atoms": (c_int, [c_int]),
    "geossl_cfconv_pairs": (c_int, [c_p, c_p, c_p, c_p, c_p, c_i64, c_int, c_int, c_int, c_p, c_p]),
    "geossl_pair_index": (c_int, [c_p, c_p, c_p, c_i64, c_p, c_int, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "geossl_cfconv_bwd_w": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_int, c_p, c_p]),
    "geossl_filter_bwd_workspace": (c_i64, [c_int, c_int]),
    "geossl_filter_bwd": (c_int, [c_p, c_p, c_i64, c_p, c_f, c_f, c_int, c_int, c_p, c_p, c_p, c_p,
                                  c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "geossl_filter_bwd_tc_workspace": (c_i64, []),
    "geossl_filter_bwd_tc": (c_int, [c_p, c_p, c_i64, c_p, c_f, c_f, c_int, c_int, c_p, c_p, c_p,
                                     c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "geossl_weight_image_bytes": (c_i64, []),
    "geossl_pack_weight": (c_int, [c_p, c_int, c_int, c_p, c_p]),
    "geossl_pack_weight_ld": (c_int, [c_p, c_int, c_int, c_int, c_p, c_p]),
    "geossl_pack_weight_pair": (c_int, [c_p, c_int, c_p, c_p]),
    "geossl_ddm_workspace_fused": (c_i64, [c_i64, c_i64]),
    "geossl_pack_weights_batched": (c_int, [c_p, c_p, c_int, c_p, c_p]),
    "geossl_linear_tc_block": (c_int, [c_p, c_i64, c_i64, c_p, c_p, c_int, c_int, c_p, c_i64, c_p, c_i64, c_p, c_i64, c_int, c_int, c_p]),
    "geossl_linear_wgrad_tc_block": (c_int, [c_p, c_i64, c_p, c_i64, c_i64, c_int, c_p, c_p, c_int, c_p, c_int, c_p]),
    "geossl_linear_tc": (c_int, [c_p, c_i64, c_p, c_p, c_int, c_p, c_p, c_p, c_int, c_p]),
    "geossl_linear_chain_tc": (c_int, [c_p, c_i64, ctypes.POINTER(ChainStage), c_int, c_int, c_int, c_p]),
    "geossl_linear_chain_ex": (c_int, [c_i64, ctypes.POINTER(ChainStageEx), c_int, c_int, c_int, c_p]),
    "geossl_linear_wgrad_tc_batch": (c_int, [ctypes.POINTER(WgradProblem), c_int, c_i64, c_p]),
    "geossl_linear_wgrad_tc_workspace": (c_i64, [c_i64]),
    "geossl_linear_wgrad_tc": (c_int, [c_p, c_p, c_i64, c_int, c_p, c_p, c_p, c_p]),
    "geossl_pair_distance": (c_int, [c_p, c_p, c_i64, c_p, c_p]),
    "geossl_ddm_workspace": (c_i64, [c_int]),
    "geossl_ddm_workspace_tc": (c_i64, [c_i64]),
    "geossl_ddm_head_fwd": (c_int, [c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_int, c_f, c_int,
                                    ctypes.POINTER(DdmPtrs), c_p, c_p, c_p]),
    "geossl_ddm_head_bwd": (c_int, [c_p, c_p, c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_int, c_f, c_int,
                                    ctypes.POINTER(DdmPtrs), c_p, c_p, c_p, c_p, ctypes.POINTER(DdmPtrs), c_p]),
    "geossl_ddm_head_fwd_tc": (c_int, [c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_p, c_p, c_int, c_f, c_int,
                                       ctypes.POINTER(DdmPtrs), c_p, c_p, c_p]),
    "geossl_ddm_head_bwd_tc": (c_int, [c_p, c_p, c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_int, c_f, c_int,
                                       ctypes.POINTER(DdmPtrs), c_p, c_p, c_p, c_p, ctypes.POINTER(DdmPtrs), c_p]),
    "geossl_ddm_head_fwd_bwd_tc": (c_int, [c_p, c_p, c_p, c_i64, c_p, c_i64, c_p, c_p, c_p, c_p, c_int, c_f, c_int,
                                           ctypes.POINTER(DdmPtrs), c_p, c_p, c_p, ctypes.POINTER(DdmPtrs), c_p]),
    "geossl_painn_edge_geometry": (c_int, [c_p, c_p, c_i64, c_i64, c_f, c_p, c_p, c_p, c_p]),
    "geossl_painn_message_fwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_p, c_p, c_p,
                                         c_p, c_p, c_p, c_i64, c_p, c_p, c_p, c_p]),
    "geossl_painn_rbf_pad": (c_int, [c_p, c_p, c_i64, c_p, c_p, c_int, c_int, c_p, c_p]),
    "geossl_painn_workspace": (c_i64, [c_int, c_int]),
    "geossl_painn_message_bwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_int, c_int, c_p, c_p, c_p,
                                         c_p, c_p, c_i64, c_i64, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p]),
    "geossl_painn_mix_pre": (c_int, [c_p, c_p, c_i64, c_int, c_f, c_p, c_p, c_p]),
    "geossl_painn_mix_pre_bwd": (c_int, [c_p, c_p, c_p, c_p, c_i64, c_int, c_p, c_p, c_p]),
    "geossl_painn_mix_post": (c_int, [c_p, c_p, c_p, c_p, c_p, c_i64, c_int, c_p, c_p, c_p]),
    "geossl_painn_mix_post_bwd": (c_int, [c_p, c_p, c_p, c_p, c_p, c_i64, c_int, c_p, c_p, c_p, c_p]),
}


def exported_symbols():
    """Names include/geossl_b200.h declares (kept in sync by tests/test_abi.py)."""
    return sorted(_SIGNATURES)


def sources():
    return sorted(glob.glob(os.path.join(_CSRC, "*.cu")))


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a into libgeossl_b200.so (in-tree)."""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(_CSRC, "*.cuh")) + glob.glob(os.path.join(_HERE, "..", "include", "*.h"))
    if not force and os.path.exists(LIB_PATH):
        newest = max(os.path.getmtime(p) for p in deps)
        if os.path.getmtime(LIB_PATH) >= newest:
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc, "-shared", *NVCC_FLAGS, "-o", LIB_PATH, *srcs]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def load():
    """Load the library (must have been built: `python -c 'import __graft_entry__ as g; g.build()'`)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            try:                                   # fresh checkout: compile in-tree (nvcc, ~10 s); there is no other path
                build()
            except Exception as exc:
                raise RuntimeError(
                    f"{LIB_PATH} is missing and could not be built ({exc}): geossl_b200 has no CPU / eager fallback. "
                    "Build it with `python -c \"import __graft_entry__ as g; g.build()\"` (needs nvcc).") from exc
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)     # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.geossl_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH} has ABI version {lib.geossl_abi_version()}, this package binds version {ABI_VERSION}: "
                               "rebuild it with `python -c \"import __graft_entry__ as g; g.build()\"`")
        _lib = lib
        return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().geossl_last_error().decode(errors="replace")
        raise RuntimeError(f"libgeossl_b200 call failed ({what} rc={rc}): {msg}")


def launch_count(reset=False):
    return int(load().geossl_launch_count(1 if reset else 0))
