"""Clock-stamp timeline of CTA 0 of the persistent / CTA-pair GEMM kernel (debug aid).
Usage (GPU box): [GENCAST_GEMM_PAIR=0] python tools/trace_gemm.py M N K [act] [f32res]"""
import ctypes, sys
import numpy as np
import torch
sys.path.insert(0, ".")
from gencast_flax_nnx_b200 import _lib, ops
M, N, K = (int(x) for x in sys.argv[1:4])
act = sys.argv[4] if len(sys.argv) > 4 else None
d = torch.device("cuda:0")
a = torch.randn(M, K, device=d).to(torch.bfloat16); w = (torch.randn(N, K, device=d) / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device=d); out = torch.empty(M, N, dtype=torch.bfloat16, device=d)
lib = _lib.load()
for _ in range(3): ops.gemm([(a, w)], out, bias=bias, act=act)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.gemm([(a, w)], out, bias=bias, act=act)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 100
print(f"M={M} N={N} K={K} act={act}: {us:.1f} us, {2*M*N*K/us/1e6:.0f} TFLOP/s")
trace = torch.zeros(8 * 512, dtype=torch.int64, device=d)
lib.gc_debug_set_gemm_trace(ctypes.c_void_p(trace.data_ptr()))
ops.gemm([(a, w)], out, bias=bias, act=act)
torch.cuda.synchronize()
lib.gc_debug_set_gemm_trace(ctypes.c_void_p(0))
t = trace.cpu().numpy().reshape(8, 512)
t0 = t[t > 0].min()
rel = lambda x: int(x - t0) if x > 0 else -1
nt = int((t[1][::4] > 0).sum())
print("tiles of CTA 0:", nt)
print("TMA issue times (first 24):", [rel(x) for x in t[0][:24]])
print("MMA per tile (start wait acc_empty, acc free, all issued):", [(rel(t[1][4*i]), rel(t[1][4*i+1]), rel(t[1][4*i+2])) for i in range(nt)])
print("EPI per tile (start wait acc_full, acc arrived, done):", [(rel(t[2][4*i]), rel(t[2][4*i+1]), rel(t[2][4*i+2])) for i in range(nt)])
print("MMA operand-ready times (first 40 k-blocks):", [rel(x) for x in t[3][:40]])
import numpy as np
full = np.array([x for x in t[3] if x > 0]); 
if len(full) > 16: print("median k-block period (clk):", float(np.median(np.diff(full))), " p90:", float(np.percentile(np.diff(full), 90)))
fine = t[4:].reshape(-1); fine = fine[fine > 0]
if len(fine):
    print("epilogue warp 2 fine stamps (bf16 staged: per 32-col chunk = start, acc in regs, staging free, stored):")
    print([rel(x) for x in fine[:64]])
