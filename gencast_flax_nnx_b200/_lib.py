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


class Mlp2(Structure):
    """Mirror of struct gc_mlp2."""
    _fields_ = [("w1", c_void_p * GC_MAX_SEGMENTS), ("k1", c_int32 * GC_MAX_SEGMENTS), ("num_segments", c_int32),
                ("b1", c_void_p), ("w2", c_void_p), ("b2", c_void_p)]


class TransformerLayer(Structure):
    """Mirror of struct gc_transformer_layer."""
    _fields_ = [("wqkv", c_void_p), ("wo", c_void_p), ("bo", c_void_p), ("w1", c_void_p), ("b1", c_void_p),
                ("w2", c_void_p), ("b2", c_void_p)]


class DenoiserModel(Structure):
    """Mirror of struct gc_denoiser_model."""
    _fields_ = [("dtype", c_int32), ("latent", c_int32), ("heads", c_int32), ("head_dim", c_int32), ("ffw_hidden", c_int32),
                ("num_layers", c_int32), ("n_out_padded", c_int32), ("reserved", c_int32),
                ("grid_embed", Mlp2), ("g2m_w1s", c_void_p), ("g2m_w2", c_void_p), ("g2m_b2", c_void_p),
                ("mesh_update", Mlp2), ("grid_update", Mlp2), ("layers", POINTER(TransformerLayer)),
                ("m2g_w1s", c_void_p), ("m2g_w1r", c_void_p), ("m2g_w2", c_void_p), ("m2g_b2", c_void_p),
                ("m2g_grid_update", Mlp2), ("output", Mlp2)]


class DenoiserGraph(Structure):
    """Mirror of struct gc_denoiser_graph."""
    _fields_ = [("grid_rows", c_int64), ("mesh_rows", c_int64), ("g2m_edges", c_int64), ("m2g_edges", c_int64),
                ("g2m_senders", c_void_p), ("g2m_receivers", c_void_p), ("g2m_row_ptr", c_void_p), ("g2m_perm", c_void_p),
                ("m2g_senders", c_void_p), ("m2g_receivers", c_void_p), ("m2g_row_ptr", c_void_p), ("m2g_perm", c_void_p),
                ("g2m_edge_ln", c_void_p), ("m2g_edge_ln", c_void_p),
                ("attention_kind", c_int32), ("max_degree", c_int32), ("num_q_tiles", c_int32), ("mask_period", c_int32),
                ("step_ptr", c_void_p), ("keys", c_void_p), ("step_mask", c_void_p), ("work", c_void_p),
                ("tile_ptr", c_void_p), ("tile_kv", c_void_p), ("tile_mask", c_void_p),
                ("nbr_ptr", c_void_p), ("nbr_idx", c_void_p)]


class SigmaContextC(Structure):
    """Mirror of struct gc_sigma_context."""
    _fields_ = [("table", c_void_p), ("g2m_w1e", c_void_p), ("g2m_b1", c_void_p), ("m2g_w1e", c_void_p), ("m2g_b1", c_void_p),
                ("g2m_base", c_void_p), ("m2g_base", c_void_p), ("g2m_base_rows", c_int64), ("m2g_base_rows", c_int64),
                ("m0", c_void_p), ("m_p", c_void_p)]


class DenoiserWorkspace(Structure):
    """Mirror of struct gc_denoiser_workspace."""
    _fields_ = [(n, c_void_p) for n in (
        "xin", "a_const", "g_h", "g_y", "g_h2", "g_y2", "g0", "g_lat", "g2", "g_p", "g_p2", "g_agg",
        "m_h", "m_y", "m_p", "m_agg", "m_out", "t_h", "t_o", "x", "t_qkv", "t_f", "e_h", "e_y", "f_out",
        "branch_stream", "fork_event", "join_event")] + [("flags", c_int32), ("reserved", c_int32)]


GC_ATTENTION_CSR, GC_ATTENTION_TILES, GC_ATTENTION_GATHER = 0, 1, 2
GC_FORWARD_FUSE_M2G = 1
GC_FORWARD_FUSE_LN = 2
GC_FORWARD_FUSE_G2M = 4
FORWARD_STRUCTS = (DenoiserModel, DenoiserGraph, SigmaContextC, DenoiserWorkspace, Mlp2, TransformerLayer)

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
    "gc_ln_cond_segment_sum_stats": (c_int32, [c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_int32, c_void_p, c_void_p,
                                               c_void_p, c_int32, c_int64, c_int64, c_int32, c_void_p]),
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
    "gc_normalize_cast": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_int32, c_int64, c_int64]),
    "gc_unnormalize_residual": (c_int32, [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int64,
                                          c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_int64]),
    "gc_edge_hidden": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                 c_int64, c_int32, c_void_p, c_int64, c_int64, c_int32]),
    "gc_linear_ln_cond": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int32,
                                    c_void_p, c_int32, c_int64, c_void_p, c_int32, c_int64, c_int32]),
    "gc_edge_mlp_sum3": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                   c_int64, c_int32, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_int32,
                                   c_int64, c_int64, c_int32]),
    "gc_edge_mlp_rows": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_int64,
                                   c_void_p, c_void_p, c_int64, c_int64, c_int32, c_void_p]),
    "gc_denoiser_forward": (c_int32, [c_void_p, POINTER(DenoiserModel), POINTER(DenoiserGraph), POINTER(SigmaContextC),
                                      POINTER(DenoiserWorkspace)]),
    "gc_sizeof_forward_structs": (c_int32, [c_int32]),
    "gc_sh_synthesis": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32]),
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
    try:
        from . import build as _build
        if (_build.CSRC / "abi.cu").exists() and not _build.is_current():
            import warnings
            warnings.warn(f"{LIB_PATH.name} is older than the sources under csrc/ (fingerprint mismatch): rebuild with "
                          "`python -m gencast_flax_nnx_b200.build`", RuntimeWarning)
    except ImportError:
        pass
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.gc_sizeof_gemm_args() != ctypes.sizeof(GemmArgs):
        raise GencastKernelError("struct gc_gemm_args layout mismatch between the library and the ctypes binding")
    for i, st in enumerate(FORWARD_STRUCTS):
        if lib.gc_sizeof_forward_structs(i) != ctypes.sizeof(st):
            raise GencastKernelError(f"struct layout mismatch between the library and the ctypes binding: {st.__name__} "
                                     f"({lib.gc_sizeof_forward_structs(i)} vs {ctypes.sizeof(st)})")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().gc_last_error()
        raise GencastKernelError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")
