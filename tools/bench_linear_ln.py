"""Times gc_linear_ln_cond (second MLP layer + LayerNorm + affine + residual in one kernel) against GEMM -> gc_ln_cond on
the GenCast 1 deg x 4 members grid-node shape (debug / profiling aid).  Usage (GPU box): python tools/bench_linear_ln.py"""
import sys

import torch

sys.path.insert(0, ".")
from gencast_flax_nnx_b200 import ops

d = torch.device("cuda:0")
g = torch.Generator(device=d).manual_seed(0)
bf = torch.bfloat16
L = 512
for rows in (4 * 65160, 4 * 10368):
    a = torch.randn(rows, L, device=d, generator=g).to(bf)
    w = (torch.randn(L, L, device=d, generator=g) / 22.6).to(bf)
    b = torch.randn(L, device=d, generator=g) * 0.1
    so = torch.cat([1 + 0.1 * torch.randn(L, device=d, generator=g), torch.randn(L, device=d, generator=g)])
    res = torch.randn(rows, L, device=d, generator=g).to(bf)
    y = torch.empty(rows, L, dtype=bf, device=d)
    out = torch.empty(rows, L, dtype=bf, device=d)
    cases = {
        "fused, no residual": lambda: ops.linear_ln_cond(a, w, b, so, out),
        "fused, bf16 residual": lambda: ops.linear_ln_cond(a, w, b, so, out, residual=res),
        "gemm + ln_cond, no residual": lambda: (ops.gemm([(a, w)], y, bias=b, static_weights=True), ops.ln_cond(y, out, so)),
        "gemm + ln_cond, bf16 residual": lambda: (ops.gemm([(a, w)], y, bias=b, static_weights=True), ops.ln_cond(y, out, so, residual=res)),
    }
    for name, fn in cases.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            fn()
        t1.record()
        torch.cuda.synchronize()
        print(f"rows {rows:7d}: {name:32s} {t0.elapsed_time(t1) / 10 * 1e3:8.1f} us")
