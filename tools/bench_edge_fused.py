"""Times gc_edge_mlp_sum3 (fused mesh2grid edge update + aggregation) against the three-kernel path on the
GenCast 1 deg x 4 members shape (debug / profiling aid).  Usage (GPU box): python tools/bench_edge_fused.py [members]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gencast_flax_nnx_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
G, V, L = 65160, 10368, 512
d = torch.device("cuda:0")
g = torch.Generator(device=d).manual_seed(0)
bf = torch.bfloat16
R, E = B * G, 3 * B * G
base = torch.randn(3 * G, L, device=d, generator=g).to(bf)
gs = torch.randn(B * V, L, device=d, generator=g).to(bf)
gr = torch.randn(R, L, device=d, generator=g).to(bf)
idx_s = (torch.randint(0, V, (E,), device=d, generator=g) + torch.arange(B, device=d).repeat_interleave(3 * G) * V).to(torch.int32)
idx_r = torch.arange(R, device=d, dtype=torch.int32).repeat_interleave(3)
w2 = (torch.randn(L, L, device=d, generator=g) / 22.6).to(bf)
b2 = torch.randn(L, device=d, generator=g) * 0.1
so = torch.cat([1 + 0.1 * torch.randn(L, device=d, generator=g), torch.randn(L, device=d, generator=g)])
out = torch.empty(R, L, dtype=bf, device=d)
e_h = torch.empty(E, L, dtype=bf, device=d)
e_y = torch.empty(E, L, dtype=bf, device=d)
rp = torch.arange(0, 3 * R + 1, 3, dtype=torch.int32, device=d)


def fused():
    # no receiver table: receiver v's row is gr[v] (base / receiver rows by TMA; GENCAST_EDGE_TMA=0 = all through registers)
    ops.edge_mlp_sum3(base, [(gs, idx_s), (gr, None)], w2, b2, so, out)


def unfused():
    ops.edge_hidden(base, [(gs, idx_s), (gr, idx_r)], e_h, act="swish")
    ops.gemm([(e_h, w2)], e_y, bias=b2, static_weights=True)
    ops.ln_cond_segment_sum(e_y, out, so, rp, None)


for name, fn in (("fused", fused), ("three kernels", unfused)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        fn()
    b.record()
    torch.cuda.synchronize()
    print(f"{name}: {a.elapsed_time(b) / 10 * 1e3:.1f} us  ({B} members, {E} edges, L={L})")

# ---- clock-stamp timeline of CTA 0 of the fused kernel (tiles blockIdx.x, + gridDim.x, ...)
import ctypes

from gencast_flax_nnx_b200 import _lib

lib = _lib.load()
trace = torch.zeros(8 * 512, dtype=torch.int64, device=d)
lib.gc_debug_set_edge_fused_trace(ctypes.c_void_p(trace.data_ptr()))
fused()
torch.cuda.synchronize()
lib.gc_debug_set_edge_fused_trace(ctypes.c_void_p(0))
tr = trace.cpu().numpy().reshape(8, 512)
t0 = tr[tr > 0].min()
rel = lambda x: int(x - t0) if x > 0 else -1
KB = L // 64
nt = min(6, int((tr[0][2::3] > 0).sum()))
print("tiles of CTA 0 shown:", nt, "(cycles since the first stamp)")
print("MMA per tile (wait acc_empty start, acc free, all issued):", [(rel(tr[0][3 * i]), rel(tr[0][3 * i + 1]), rel(tr[0][3 * i + 2])) for i in range(nt)])
print("MMA: k-block operands ready at:", [[rel(tr[1][i * KB + k]) for k in range(KB)] for i in range(nt)])
print("producer warp 2 per k-block (wait TMA start, landed, stage signalled):")
for i in range(nt):
    print("   ", [(rel(tr[2][3 * (i * KB + k)]), rel(tr[2][3 * (i * KB + k) + 1]), rel(tr[2][3 * (i * KB + k) + 2])) for k in range(KB)])
print("epilogue warp 10 per tile (wait acc start, acc arrived, statistics done, tile stored):",
      [tuple(rel(tr[3][4 * i + j]) for j in range(4)) for i in range(nt)])
print("epilogue warp 10, third tile, pass 2 per 32-column chunk (start, accumulator in registers, normalised + summed, stored):",
      [tuple(rel(tr[4][4 * i + j]) for j in range(4)) for i in range(L // 64)])
