"""ctypes binding of the C ABI declared in include/gencast_b200.h.

There is no CPU fallback: if the shared library has not been built
(`python -m gencast_flax_nnx_b200.build` or `__graft_entry__.build()`), loading
raises.  Every launcher is enqueue-only on the CUDA stream passed in.
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_void_p
from pathlib import Path

GC_F32 = 0
GC_BF16 = 1
GC_ACT_NONE = 0
GC_ACT_SWISH = 1
GC_ACT_GELU_TANH = 2
GC_MAX_SEGMENTS = 3
GC_GEMM_STATIC_WEIGHTS = 1

LIB_PATH = Path(__file__).resolve().parent / "libgencast_b200.so"


class GemmArgs(Structure):
    """Mirror of struct gc_gemm_args."""
    _fields_ = [
        ("a", c_void_p * GC_MAX_SEGMENTS),
        ("w", c_void_p * GC_MAX_SEGMENTS),
        ("lda", c_int64 * GC_MAX_SEGMENTS),
        ("ldw", c_int64 * GC_MAX_SEGMENTS),
        ("k", c_int32 * GC_MAX_SEGMENTS),
        ("num_segments", c_int32),
        ("m", c_int64),
        ("n", c_int32),
        ("dtype", c_int32),
        ("bias", c_void_p),
        ("alpha_dev", c_void_p),
        ("addend", c_void_p),
        ("ld_addend", c_int64),
        ("addend_dtype", c_int32),
        ("gather_dtype", c_int32),
        ("gather_src", c_void_p * 2),
        ("gather_idx", c_void_p * 2),
        ("ld_gather", c_int64 * 2),
        ("act", c_int32),
        ("res_dtype", c_int32),
        ("residual", c_void_p),
        ("ld_res", c_int64),
        ("out", c_void_p),
        ("ldo", c_int64),
        ("out_dtype", c_int32),
        ("flags", c_int32),
    ]


# name -> (restype, argtypes); also the list of symbols tests check for.
SIGNATURES = {
    "gc_last_error": (c_char_p, []),
    "gc_abi_version": (c_int32, []),
    "gc_device_supports_tcgen05": (c_int32, []),
    "gc_gemm": (c_int32, [c_void_p, POINTER(GemmArgs)]),
    "gc_sizeof_gemm_args": (c_int32, []),
    "gc_ln_cond": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_int32, c_void_p, c_int32, c_int64,
                             c_void_p, c_int32, c_int64, c_int64, c_int32]),
    "gc_ln_cond_segment_sum": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_int32, c_void_p, c_void_p,
                                         c_void_p, c_int32, c_int64, c_int64, c_int32]),
    "gc_khop_attention": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_int32, c_void_p,
                                    c_int64, c_int64, c_int32, c_int32]),
    "gc_khop_attention_tiles": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_int64, c_int64, c_int32, c_int32]),
    "gc_khop_attention_gather": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                           c_int32, c_void_p, c_int64, c_int64, c_int32, c_int32]),
    "gc_cond_tables": (c_int32, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                 c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p]),
    "gc_fold_affine_into_linear": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p,
                                             c_int64, c_void_p, c_int32, c_int32]),
    "gc_dpm_update": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                c_void_p, c_int32, c_int64, c_int64, c_int32]),
    "gc_cast_pad": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p, c_int32, c_int64, c_int32,
                              c_void_p, c_int64]),
    "gc_select_columns": (c_int32, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                    c_int64, c_int64, c_int32]),
    "gc_edge_hidden": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_int64, c_int32, c_void_p, c_int64, c_int64, c_int32]),
    "gc_edge_mlp_sum3": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                   c_int64, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_int32,
                                   c_int64, c_int64, c_int32]),
    "gc_fair_crps": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_int64]),
    "gc_column_sums": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p]),
    "gc_ensemble_accumulate": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64]),
}

_lib = None


class GencastKernelError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise GencastKernelError(
            f"{LIB_PATH} is missing: build the CUDA kernels first (python -m gencast_flax_nnx_b200.build). "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gc_sizeof_gemm_args() != ctypes.sizeof(GemmArgs):
        raise GencastKernelError("struct gc_gemm_args layout mismatch between the library and the ctypes binding")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gc_last_error()
        raise GencastKernelError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
