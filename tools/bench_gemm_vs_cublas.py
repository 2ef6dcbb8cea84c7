"""Sustained throughput of gc_gemm against torch.matmul (cuBLAS) on the GEMM shapes of one GenCast 1 deg x 4 members
evaluation, each run back to back for about a second (so both sit under the same power cap).  Debug / profiling aid.
Usage (GPU box): python tools/bench_gemm_vs_cublas.py"""
import sys
import time

import torch

sys.path.insert(0, ".")
from gencast_flax_nnx_b200 import ops

d = torch.device("cuda:0")
bf = torch.bfloat16
shapes = [("QKV", 41472, 1536, 512), ("out-proj", 41472, 512, 512), ("FFW-in", 41472, 2048, 512), ("FFW-out", 41472, 512, 2048),
          ("grid MLP", 260640, 512, 512), ("g2m edge 2nd", 407568, 512, 512)]
for name, m, n, k in shapes:
    a = torch.randn(m, k, device=d).to(bf)
    w = (torch.randn(n, k, device=d) / 22.6).to(bf)
    out = torch.empty(m, n, dtype=bf, device=d)
    res = {}
    for impl, fn in (("gc_gemm", lambda: ops.gemm([(a, w)], out, static_weights=True)), ("cuBLAS", lambda: torch.matmul(a, w.t(), out=out))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        iters = 0
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        start = time.perf_counter()
        while time.perf_counter() - start < 1.0:
            for _ in range(50):
                fn()
            iters += 50
            torch.cuda.synchronize()
        t1.record()
        torch.cuda.synchronize()
        us = t0.elapsed_time(t1) / iters * 1e3
        res[impl] = (us, 2.0 * m * n * k / us / 1e6)
    print(f"{name:14s} m={m:6d} n={n:4d} k={k:4d}: gc_gemm {res['gc_gemm'][0]:7.1f} us {res['gc_gemm'][1]:6.0f} TF/s | cuBLAS {res['cuBLAS'][0]:7.1f} us "
          f"{res['cuBLAS'][1]:6.0f} TF/s | ratio {res['gc_gemm'][1] / res['cuBLAS'][1]:.2f}")
