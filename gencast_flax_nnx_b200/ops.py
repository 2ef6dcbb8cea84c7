"""Torch-tensor front end of the C ABI (device memory and streams are torch's; the
compute is the library's).  Every function enqueues on torch's current stream and
returns immediately; outputs are caller-allocated unless stated.

Reference operators these replace are cited in include/gencast_b200.h.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import GC_ACT_GELU_TANH, GC_ACT_NONE, GC_ACT_SWISH, GC_BF16, GC_F32, GemmArgs

# Optional per-launch timing (bench.py's roofline leg): when a recorder is installed every
# launcher is bracketed by CUDA events on the launching stream and reports its
# algorithmic FLOPs / bytes.  Never active during graph capture or normal runs.
_RECORDER = None


def set_recorder(rec) -> None:
    global _RECORDER
    _RECORDER = rec


class Recorder:
    def __init__(self):
        self.items = []      # (name, start_event, end_event, flops, bytes)

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, a, b, fl, by in self.items:
            d = agg.setdefault(name, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d["launches"] += 1; d["ms"] += a.elapsed_time(b); d["flops"] += fl; d["bytes"] += by
        return agg


def _recorded(name, cost):
    def deco(fn):
        def wrapper(*args, **kw):
            rec = _RECORDER
            if rec is None:
                return fn(*args, **kw)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn(*args, **kw)
            b.record()
            fl, by = cost(*args, **kw)
            rec.items.append((name(*args, **kw) if callable(name) else name, a, b, fl, by))
            return out
        wrapper.__doc__ = fn.__doc__
        wrapper.__name__ = fn.__name__
        return wrapper
    return deco


def _nbytes(*ts):
    return float(sum(t.numel() * t.element_size() for t in ts if t is not None))


def _gemm_cost(segments, out, *, bias=None, act=None, addend=None, gathers=(), residual=None, alpha=None,
               static_weights=False):
    m, n = out.shape
    flops = 2.0 * m * n * sum(a.shape[1] for a, _ in segments)
    by = sum(_nbytes(a, w) for a, w in segments) + _nbytes(out, addend, residual)
    by += sum(m * n * src.element_size() + 4.0 * m for src, _ in gathers)
    return flops, by


def _gemm_name(segments, out, **kw):
    return "gemm_bf16_tcgen05" if segments[0][0].dtype == torch.bfloat16 else "gemm_f32_ffma"


ACT = {None: GC_ACT_NONE, "none": GC_ACT_NONE, "swish": GC_ACT_SWISH, "gelu_tanh": GC_ACT_GELU_TANH}


def _dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return GC_F32
    if t.dtype == torch.bfloat16:
        return GC_BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _row_major(t: torch.Tensor, what: str) -> int:
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{what}: expected a 2-D tensor with unit inner stride, got shape {tuple(t.shape)} "
                         f"strides {t.stride()}")
    if not t.is_cuda:
        raise ValueError(f"{what}: tensor is not on a CUDA device (there is no CPU path)")
    return t.stride(0)


@_recorded(_gemm_name, _gemm_cost)
def gemm(segments: Sequence[Tuple[torch.Tensor, torch.Tensor]], out: torch.Tensor, *,
         bias: Optional[torch.Tensor] = None, act: Optional[str] = None,
         addend: Optional[torch.Tensor] = None,
         gathers: Sequence[Tuple[torch.Tensor, torch.Tensor]] = (),
         residual: Optional[torch.Tensor] = None, alpha: Optional[torch.Tensor] = None,
         static_weights: bool = False) -> torch.Tensor:
    """out = act(alpha * sum_s A_s @ W_s^T + bias + addend + sum_j G_j[idx_j]) + residual.

    segments: [(A_s [m, k_s], W_s [n, k_s])], all of one dtype (bf16 -> tcgen05, f32 -> FFMA).
    """
    lib = _lib.load()
    args = GemmArgs()
    m, n = out.shape
    for s, (a, w) in enumerate(segments):
        if a.dtype != w.dtype or a.dtype != segments[0][0].dtype:
            raise TypeError("gemm: all operands must share one dtype")
        if a.shape[0] != m or w.shape[0] != n or a.shape[1] != w.shape[1]:
            raise ValueError(f"gemm: segment {s} shapes {tuple(a.shape)} x {tuple(w.shape)} do not match out {tuple(out.shape)}")
        args.a[s] = a.data_ptr(); args.w[s] = w.data_ptr()
        args.lda[s] = _row_major(a, "A"); args.ldw[s] = _row_major(w, "W")
        args.k[s] = a.shape[1]
    args.num_segments = len(segments)
    args.m = m; args.n = n
    args.dtype = _dt(segments[0][0])
    args.bias = _p(bias)
    args.alpha_dev = _p(alpha)
    if addend is not None:
        args.addend = addend.data_ptr(); args.ld_addend = _row_major(addend, "addend"); args.addend_dtype = _dt(addend)
    for j, (src, idx) in enumerate(gathers):
        if idx.dtype != torch.int32 or idx.numel() != m:
            raise ValueError("gemm: gather index must be int32 [m]")
        args.gather_src[j] = src.data_ptr(); args.gather_idx[j] = idx.data_ptr()
        args.ld_gather[j] = _row_major(src, "gather source")
        args.gather_dtype = _dt(src)
        if j and _dt(src) != _dt(gathers[0][0]):
            raise TypeError("gemm: gather sources must share one dtype")
    args.act = ACT[act]
    if residual is not None:
        args.residual = residual.data_ptr(); args.ld_res = _row_major(residual, "residual"); args.res_dtype = _dt(residual)
    args.out = out.data_ptr(); args.ldo = _row_major(out, "out"); args.out_dtype = _dt(out)
    # static_weights: the W matrices are not produced by earlier work on this stream (model weights), so the
    # kernel may fetch their first tiles while the preceding kernel is still draining
    args.flags = _lib.GC_GEMM_STATIC_WEIGHTS if static_weights else 0
    _lib.check(lib.gc_gemm(_stream(), ctypes.byref(args)), "gc_gemm")
    return out


@_recorded("ln_cond", lambda x, out, so, **kw: (0.0, _nbytes(x, out, kw.get("residual"))))
def ln_cond(x: torch.Tensor, out: torch.Tensor, scale_offset: Optional[torch.Tensor], *,
            layer_norm: bool = True, residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    lib = _lib.load()
    rows, cols = x.shape
    _lib.check(lib.gc_ln_cond(_stream(), x.data_ptr(), _dt(x), _row_major(x, "x"), _p(scale_offset), int(layer_norm),
                              _p(residual), _dt(residual) if residual is not None else 0,
                              _row_major(residual, "residual") if residual is not None else 0,
                              out.data_ptr(), _dt(out), _row_major(out, "out"), rows, cols), "gc_ln_cond")
    return out


@_recorded(lambda y, out, so, rp, perm, **kw: "ln_cond_segment_sum", lambda y, out, so, rp, perm, **kw: (0.0, _nbytes(y, out, rp, perm)))
def ln_cond_segment_sum(y: torch.Tensor, out: torch.Tensor, scale_offset: Optional[torch.Tensor],
                        row_ptr: torch.Tensor, edge_perm: Optional[torch.Tensor], *, layer_norm: bool = True,
                        irregular: bool = False, row_stats: Optional[torch.Tensor] = None):
    """`irregular`: hint that segment lengths vary widely (grid2mesh: 3 .. 594 edges per mesh node); selects the
    higher-occupancy kernel variant (GC_SEGSUM_IRREGULAR).  `row_stats`: fp32 [rows of y, 4] written by edge_mlp_rows
    (gc_ln_cond_segment_sum_stats)."""
    lib = _lib.load()
    nseg, cols = out.shape
    if row_ptr.dtype != torch.int32 or row_ptr.numel() != nseg + 1:
        raise ValueError("row_ptr must be int32 [num_segments + 1]")
    if row_stats is not None and (row_stats.dtype != torch.float32 or row_stats.shape != (y.shape[0], 4) or not row_stats.is_contiguous()):
        raise ValueError("row_stats must be contiguous fp32 [rows of y, 4]")
    _lib.check(lib.gc_ln_cond_segment_sum_stats(_stream(), y.data_ptr(), _dt(y), _row_major(y, "y"), _p(scale_offset),
                                                int(layer_norm) | (2 if irregular else 0), row_ptr.data_ptr(), _p(edge_perm),
                                                out.data_ptr(), _dt(out), _row_major(out, "out"), nseg, cols, _p(row_stats)),
               "gc_ln_cond_segment_sum")
    return out


@_recorded("khop_attention", lambda qkv, out, ptr, idx, heads, head_dim, *a: (4.0 * idx.numel() * heads * head_dim, _nbytes(qkv, out, ptr, idx)))
def khop_attention(qkv: torch.Tensor, out: torch.Tensor, nbr_ptr: torch.Tensor, nbr_idx: torch.Tensor,
                   heads: int, head_dim: int, max_degree: int = 0) -> torch.Tensor:
    lib = _lib.load()
    if out.dtype != qkv.dtype:
        raise TypeError("khop_attention: out must have the dtype of qkv")
    _lib.check(lib.gc_khop_attention(_stream(), qkv.data_ptr(), _dt(qkv), _row_major(qkv, "qkv"), nbr_ptr.data_ptr(),
                                     nbr_idx.data_ptr(), max_degree, out.data_ptr(), _row_major(out, "out"),
                                     qkv.shape[0], heads, head_dim), "gc_khop_attention")
    return out


@_recorded("khop_attention_tc", lambda qkv, out, tp, tk, tm, heads, head_dim, nnz=0: (4.0 * nnz * heads * head_dim, _nbytes(qkv, out, tm)))
def khop_attention_tiles(qkv: torch.Tensor, out: torch.Tensor, tile_ptr: torch.Tensor, tile_kv: torch.Tensor,
                         tile_mask: torch.Tensor, heads: int, head_dim: int, nnz: int = 0) -> torch.Tensor:
    """Tensor-core k-hop attention over a block-sparse tile list (bf16).  `nnz` (pattern size) is only
    used by the timing recorder to report algorithmic FLOPs."""
    lib = _lib.load()
    if qkv.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        raise TypeError("khop_attention_tiles: bf16 only")
    _lib.check(lib.gc_khop_attention_tiles(_stream(), qkv.data_ptr(), _row_major(qkv, "qkv"), tile_ptr.data_ptr(),
                                           tile_kv.data_ptr(), tile_mask.data_ptr(), out.data_ptr(),
                                           _row_major(out, "out"), qkv.shape[0], heads, head_dim),
               "gc_khop_attention_tiles")
    return out


@_recorded("khop_attention_gather", lambda qkv, out, sp, keys, mask, work, heads, head_dim, mask_period=0, nnz=0: (4.0 * nnz * heads * head_dim, _nbytes(qkv, out, mask, keys)))
def khop_attention_gather(qkv: torch.Tensor, out: torch.Tensor, step_ptr: torch.Tensor, keys: torch.Tensor,
                          mask: torch.Tensor, work: torch.Tensor, heads: int, head_dim: int, mask_period: int = 0,
                          nnz: int = 0) -> torch.Tensor:
    """Tensor-core k-hop attention over per-query-tile compacted key lists (graph.khop_compact_steps); bf16.
    `nnz` (pattern size) is only used by the timing recorder to report algorithmic FLOPs."""
    lib = _lib.load()
    if qkv.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        raise TypeError("khop_attention_gather: bf16 only")
    for t, name in ((step_ptr, "step_ptr"), (keys, "keys"), (mask, "mask"), (work, "work")):
        if t.dtype != torch.int32 or not t.is_cuda or not t.is_contiguous():
            raise TypeError(f"khop_attention_gather: {name} must be a contiguous int32 CUDA tensor")
    nq = work.numel()
    if step_ptr.numel() != nq + 1:
        raise ValueError("khop_attention_gather: step_ptr must have one entry per query tile plus one")
    _lib.check(lib.gc_khop_attention_gather(_stream(), qkv.data_ptr(), _row_major(qkv, "qkv"), step_ptr.data_ptr(),
                                            keys.data_ptr(), mask.data_ptr(), work.data_ptr(), nq, int(mask_period),
                                            out.data_ptr(), _row_major(out, "out"), qkv.shape[0], heads, head_dim),
               "gc_khop_attention_gather")
    return out


def cond_tables(sigma: torch.Tensor, w0, b0, w1, b1, base_period: float, num_frequencies: int,
                wc: torch.Tensor, bc: torch.Tensor, table: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    layers, _, two_w = wc.shape
    _lib.check(lib.gc_cond_tables(_stream(), sigma.data_ptr(), sigma.numel(), w0.data_ptr(), b0.data_ptr(),
                                  w1.data_ptr(), b1.data_ptr(), float(base_period), num_frequencies, wc.data_ptr(),
                                  bc.data_ptr(), layers, two_w // 2, table.data_ptr()), "gc_cond_tables")
    return table


def fold_affine_into_linear(w: torch.Tensor, bias: Optional[torch.Tensor], scale_offset: torch.Tensor,
                            w_out: torch.Tensor, bias_out: torch.Tensor):
    lib = _lib.load()
    n, k = w.shape
    _lib.check(lib.gc_fold_affine_into_linear(_stream(), w.data_ptr(), _dt(w), _row_major(w, "w"), _p(bias),
                                              scale_offset.data_ptr(), w_out.data_ptr(), _row_major(w_out, "w_out"),
                                              bias_out.data_ptr(), n, k), "gc_fold_affine_into_linear")
    return w_out, bias_out


@_recorded("dpm_update", lambda f, x_cur, x_base, sched, x_out, xin_out, cols: (0.0, 4.0 * x_cur.shape[0] * cols * 4 + (0 if xin_out is None else x_cur.shape[0] * cols * xin_out.element_size())))
def dpm_update(f: torch.Tensor, x_cur: torch.Tensor, x_base: torch.Tensor, sched: torch.Tensor,
               x_out: torch.Tensor, xin_out: Optional[torch.Tensor], cols: int):
    lib = _lib.load()
    rows = x_cur.shape[0]
    _lib.check(lib.gc_dpm_update(_stream(), f.data_ptr(), _row_major(f, "f"), x_cur.data_ptr(), x_base.data_ptr(),
                                 _row_major(x_cur, "x"), sched.data_ptr(), x_out.data_ptr(), _p(xin_out),
                                 _dt(xin_out) if xin_out is not None else 0,
                                 _row_major(xin_out, "xin") if xin_out is not None else 0, rows, cols), "gc_dpm_update")


def cast_pad(src: torch.Tensor, dst: torch.Tensor, cols_src: Optional[int] = None,
             scale: Optional[torch.Tensor] = None):
    lib = _lib.load()
    rows = src.shape[0]
    cs = src.shape[1] if cols_src is None else cols_src
    _lib.check(lib.gc_cast_pad(_stream(), src.data_ptr(), _dt(src), _row_major(src, "src"), cs, dst.data_ptr(),
                               _dt(dst), _row_major(dst, "dst"), dst.shape[1], _p(scale), rows), "gc_cast_pad")
    return dst


def normalize_cast(src: torch.Tensor, dst: torch.Tensor, loc=None, scale=None, fill_pre=None, fill_post=None) -> torch.Tensor:
    """dst[:, c] = cast(fill_post((fill_pre(src[:, c]) - loc[c]) / scale[c])) (gc_normalize_cast); src fp32 [rows, C]."""
    lib = _lib.load()
    if src.dtype != torch.float32:
        raise TypeError("normalize_cast: fp32 source expected")
    rows, cols = src.shape
    for v in (loc, scale, fill_pre, fill_post):
        if v is not None and (v.dtype != torch.float32 or v.numel() != cols):
            raise ValueError("normalize_cast: per-channel vectors must be fp32 [cols]")
    _lib.check(lib.gc_normalize_cast(_stream(), src.data_ptr(), _row_major(src, "src"), cols, _p(loc), _p(scale), _p(fill_pre),
                                     _p(fill_post), dst.data_ptr(), _dt(dst), _row_major(dst, "dst"), rows), "gc_normalize_cast")
    return dst


def unnormalize_residual(pred: torch.Tensor, cols: int, out: torch.Tensor, scale=None, loc=None, window=None, res_col=None,
                         res_fill=None, nan_cols=None) -> torch.Tensor:
    """out[:, c] = pred[:, c] * scale[c] + loc[c] + window[:, res_col[c]]; NaN where listed input columns are NaN
    (gc_unnormalize_residual)."""
    lib = _lib.load()
    rows = pred.shape[0]
    npc = 0 if nan_cols is None else nan_cols.shape[1]
    _lib.check(lib.gc_unnormalize_residual(_stream(), pred.data_ptr(), _row_major(pred, "pred"), cols, _p(scale), _p(loc),
                                           _p(window), _row_major(window, "window") if window is not None else 0,
                                           _p(res_col), _p(res_fill), _p(nan_cols), npc, out.data_ptr(), _row_major(out, "out"),
                                           rows), "gc_unnormalize_residual")
    return out


def select_columns(sources: Sequence[Optional[torch.Tensor]], table: torch.Tensor, out: torch.Tensor):
    """out[r, j] = sources[table[j] >> 24][r, table[j] & 0xffffff] (fp32 matrices; the rollout's window update)."""
    lib = _lib.load()
    src = list(sources) + [None] * (3 - len(sources))
    if table.dtype != torch.int32 or table.numel() != out.shape[1]:
        raise ValueError("select_columns: table must be int32 [out columns]")
    for t in src:
        if t is not None and (t.dtype != torch.float32 or t.shape[0] != out.shape[0]):
            raise TypeError("select_columns: sources must be fp32 with the rows of out")
    if out.dtype != torch.float32:
        raise TypeError("select_columns: out must be fp32")
    ld = [(_row_major(t, "source") if t is not None else 0) for t in src]
    _lib.check(lib.gc_select_columns(_stream(), _p(src[0]), ld[0], _p(src[1]), ld[1], _p(src[2]), ld[2], table.data_ptr(),
                                     out.data_ptr(), _row_major(out, "out"), out.shape[0], out.shape[1]), "gc_select_columns")
    return out


@_recorded("edge_hidden", lambda base, gathers, out, **kw: (0.0, _nbytes(base, out) + sum(out.shape[0] * out.shape[1] * 2.0 + 4.0 * out.shape[0] for _ in gathers)))
def edge_hidden(base: torch.Tensor, gathers: Sequence[Tuple[torch.Tensor, torch.Tensor]], out: torch.Tensor, *,
                act: Optional[str] = "swish") -> torch.Tensor:
    """out[e] = act(base[e % len(base)] + sum_j gathers[j][0][gathers[j][1][e]]) (bf16); see gc_edge_hidden."""
    lib = _lib.load()
    if base.dtype != torch.bfloat16 or out.dtype != torch.bfloat16 or any(g.dtype != torch.bfloat16 for g, _ in gathers):
        raise TypeError("edge_hidden: bf16 tensors expected")
    if not 1 <= len(gathers) <= 2 or any(i.dtype != torch.int32 or i.numel() != out.shape[0] for _, i in gathers):
        raise ValueError("edge_hidden: one or two (table, int32 index [rows]) pairs expected")
    g = list(gathers) + [(None, None)]
    _lib.check(lib.gc_edge_hidden(_stream(), base.data_ptr(), _row_major(base, "base"), base.shape[0],
                                  g[0][0].data_ptr(), g[0][1].data_ptr(), _row_major(g[0][0], "gather source"),
                                  _p(g[1][0]), _p(g[1][1]), _row_major(g[1][0], "gather source") if g[1][0] is not None else 0,
                                  ACT[act], out.data_ptr(), _row_major(out, "out"), out.shape[0], out.shape[1]), "gc_edge_hidden")
    return out


@_recorded("linear_ln_cond", lambda a, w, bias, so, out, **kw: (2.0 * a.shape[0] * w.shape[0] * w.shape[1], _nbytes(a, w, out, kw.get("residual"))))
def linear_ln_cond(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], scale_offset: Optional[torch.Tensor],
                   out: torch.Tensor, *, residual: Optional[torch.Tensor] = None, layer_norm: bool = True) -> torch.Tensor:
    """out = LN(a @ w^T + bias) * scale + offset (+ residual) in one kernel (gc_linear_ln_cond); a, w bf16."""
    lib = _lib.load()
    if a.dtype != torch.bfloat16 or w.dtype != torch.bfloat16:
        raise TypeError("linear_ln_cond: bf16 operands expected")
    rows, cols = out.shape
    if a.shape != (rows, cols) or w.shape != (cols, cols):
        raise ValueError("linear_ln_cond: a must be [rows, cols] and w [cols, cols]")
    _lib.check(lib.gc_linear_ln_cond(_stream(), a.data_ptr(), _row_major(a, "a"), rows, w.data_ptr(), _row_major(w, "w"),
                                     _p(bias), _p(scale_offset), int(layer_norm), _p(residual),
                                     _dt(residual) if residual is not None else 0,
                                     _row_major(residual, "residual") if residual is not None else 0, out.data_ptr(), _dt(out),
                                     _row_major(out, "out"), cols), "gc_linear_ln_cond")
    return out


def _edge_fused_cost(base, gathers, w2, b2, so, out, **kw):
    e = 3 * out.shape[0]
    cols = out.shape[1]
    by = _nbytes(out, w2) + base.shape[0] * cols * 2.0 + sum(e * cols * 2.0 + 4.0 * e for _ in gathers)
    return 2.0 * e * cols * cols, by


@_recorded("edge_mlp_sum3", _edge_fused_cost)
def edge_mlp_sum3(base: torch.Tensor, gathers: Sequence[Tuple[torch.Tensor, torch.Tensor]], w2: torch.Tensor,
                  b2: Optional[torch.Tensor], scale_offset: Optional[torch.Tensor], out: torch.Tensor, *,
                  act: Optional[str] = "swish", layer_norm: bool = True) -> torch.Tensor:
    """out[v] = scale * sum_{e in 3v..3v+2} LN(act(base[e % len(base)] + gs[idx_s[e]] + gr[idx_r[e]]) @ w2^T + b2) + 3 * offset
    in one kernel (see gc_edge_mlp_sum3): the fused edge update + aggregation of a degree-3, receiver-major graph.
    idx_r=None: receiver v's row is gr[v] (then base / gr rows are moved by TMA when len(base) % 30 == 0)."""
    lib = _lib.load()
    (gs, idx_s), (gr, idx_r) = gathers
    if any(t.dtype != torch.bfloat16 for t in (base, gs, gr, w2)):
        raise TypeError("edge_mlp_sum3: bf16 operands expected")
    nrecv, cols = out.shape
    if idx_s.dtype != torch.int32 or idx_s.numel() != 3 * nrecv or (
            idx_r is not None and (idx_r.dtype != torch.int32 or idx_r.numel() != 3 * nrecv)):
        raise ValueError("edge_mlp_sum3: index tables must be int32 [3 * receivers]")
    if idx_r is None and gr.shape[0] < nrecv:
        raise ValueError("edge_mlp_sum3: idx_r=None means gr[v] is receiver v's row; gr has too few rows")
    if w2.shape != (cols, cols):
        raise ValueError("edge_mlp_sum3: w2 must be [cols, cols]")
    _lib.check(lib.gc_edge_mlp_sum3(_stream(), base.data_ptr(), _row_major(base, "base"), base.shape[0],
                                    gs.data_ptr(), idx_s.data_ptr(), _row_major(gs, "gs"),
                                    gr.data_ptr(), _p(idx_r), _row_major(gr, "gr"), ACT[act],
                                    w2.data_ptr(), _row_major(w2, "w2"), _p(b2), _p(scale_offset), int(layer_norm),
                                    out.data_ptr(), _dt(out), _row_major(out, "out"), nrecv, cols), "gc_edge_mlp_sum3")
    return out


def _edge_rows_cost(base, gather, w2, b2, out, **kw):  # noqa: row_stats adds 16 bytes per row
    rows, cols = out.shape
    # FLOPs of the second layer; bytes: the base table once, one gathered row and one index per edge, the rows written
    return 2.0 * rows * cols * cols, _nbytes(base, out, w2) + rows * cols * 2.0 + 4.0 * rows


@_recorded("edge_mlp_rows", _edge_rows_cost)
def edge_mlp_rows(base: torch.Tensor, gather: Tuple[torch.Tensor, torch.Tensor], w2: torch.Tensor, b2: Optional[torch.Tensor],
                  out: torch.Tensor, *, act: Optional[str] = "swish", row_stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[e] = act(base[e % len(base)] + gs[idx_s[e]]) @ w2^T + b2 in one kernel (see gc_edge_mlp_rows).  `row_stats`: optional
    fp32 [rows, 4] output, per row {sum, sum of squares} of the two column halves (for ln_cond_segment_sum)."""
    lib = _lib.load()
    gs, idx_s = gather
    if any(t.dtype != torch.bfloat16 for t in (base, gs, w2, out)):
        raise TypeError("edge_mlp_rows: bf16 operands expected")
    rows, cols = out.shape
    if idx_s.dtype != torch.int32 or idx_s.numel() != rows or rows % base.shape[0] != 0:
        raise ValueError("edge_mlp_rows: idx_s must be int32 [rows], rows a multiple of len(base)")
    if w2.shape != (cols, cols):
        raise ValueError("edge_mlp_rows: w2 must be [cols, cols]")
    _lib.check(lib.gc_edge_mlp_rows(_stream(), base.data_ptr(), _row_major(base, "base"), base.shape[0], gs.data_ptr(),
                                    idx_s.data_ptr(), _row_major(gs, "gs"), ACT[act], w2.data_ptr(), _row_major(w2, "w2"),
                                    _p(b2), out.data_ptr(), _row_major(out, "out"), rows, cols, _p(row_stats)), "gc_edge_mlp_rows")
    return out


def denoiser_forward(model, graph, sigma, workspace) -> None:
    """One whole network evaluation through gc_denoiser_forward (descriptors: _lib.DenoiserModel, DenoiserGraph,
    SigmaContextC, DenoiserWorkspace, filled by the engine)."""
    lib = _lib.load()
    _lib.check(lib.gc_denoiser_forward(_stream(), ctypes.byref(model), ctypes.byref(graph), ctypes.byref(sigma),
                                       ctypes.byref(workspace)), "gc_denoiser_forward")


def sh_synthesis(coef: torch.Tensor, table: torch.Tensor, spec: torch.Tensor, out: torch.Tensor, members: int,
                 channels: int, n_lat: int, n_lon: int) -> torch.Tensor:
    """Spherical-harmonic synthesis of random coefficients into [members * n_lat * n_lon, channels] noise (gc_sh_synthesis)."""
    lib = _lib.load()
    L = table.shape[0]
    for t in (coef, table, spec, out):
        if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
            raise TypeError("sh_synthesis: contiguous fp32 CUDA tensors expected")
    if coef.shape != (2, L, members * channels, L) or table.shape != (L, L, n_lat):
        raise ValueError("sh_synthesis: coef must be [2, L, members * channels, L] and table [L, L, n_lat]")
    if spec.numel() < members * n_lat * channels * L * 2 or out.shape != (members * n_lat * n_lon, channels):
        raise ValueError("sh_synthesis: scratch / output size")
    _lib.check(lib.gc_sh_synthesis(_stream(), coef.data_ptr(), table.data_ptr(), spec.data_ptr(), out.data_ptr(), L, members,
                                   channels, n_lat, n_lon), "gc_sh_synthesis")
    return out


def fair_crps(members: torch.Tensor, truth: torch.Tensor, weights: Optional[torch.Tensor], channels: int) -> torch.Tensor:
    """members [M, n] fp32, truth [n], weights [n / channels] or None -> weighted fair CRPS per point, [n] fp32."""
    lib = _lib.load()
    M, n = members.shape
    if members.dtype != torch.float32 or truth.dtype != torch.float32 or truth.numel() != n:
        raise TypeError("fair_crps: members [M, n] and truth [n] must be fp32")
    if weights is not None and (weights.dtype != torch.float32 or weights.numel() * channels != n):
        raise TypeError("fair_crps: weights must be fp32 [n / channels]")
    out = torch.empty(n, dtype=torch.float32, device=members.device)
    _lib.check(lib.gc_fair_crps(_stream(), members.data_ptr(), _row_major(members, "members"), M, truth.data_ptr(),
                                _p(weights), channels, out.data_ptr(), n), "gc_fair_crps")
    return out


def column_sums(x: torch.Tensor) -> torch.Tensor:
    """[rows, cols] fp32 -> [cols] sums over rows in a fixed order."""
    lib = _lib.load()
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise TypeError("column_sums: contiguous fp32 matrix expected")
    out = torch.empty(x.shape[1], dtype=torch.float32, device=x.device)
    _lib.check(lib.gc_column_sums(_stream(), x.data_ptr(), x.shape[0], x.shape[1], out.data_ptr()), "gc_column_sums")
    return out


def ensemble_accumulate(x: torch.Tensor, total: torch.Tensor, total_sq: torch.Tensor):
    lib = _lib.load()
    _lib.check(lib.gc_ensemble_accumulate(_stream(), x.data_ptr(), total.data_ptr(), total_sq.data_ptr(), x.numel()),
               "gc_ensemble_accumulate")
