"""Times gc_edge_mlp_rows (grid2mesh edge MLP: tabulated first layer + gather + second layer in one kernel) against
gc_edge_hidden -> gc_gemm at the GenCast 1 deg x 4 members shape.  Usage (GPU box): python tools/bench_edge_rows.py [members]"""
import sys

import torch

sys.path.insert(0, ".")
from gencast_flax_nnx_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
G, E1, L = 65160, 101892, 512
d = torch.device("cuda:0")
g = torch.Generator(device=d).manual_seed(0)
bf = torch.bfloat16
E = B * E1
base = torch.randn(E1, L, device=d, generator=g).to(bf)
gs = torch.randn(B * G, L, device=d, generator=g).to(bf)
idx_s = (torch.randint(0, G, (E,), device=d, generator=g) + torch.arange(B, device=d).repeat_interleave(E1) * G).to(torch.int32)
w2 = (torch.randn(L, L, device=d, generator=g) / 22.6).to(bf)
b2 = torch.randn(L, device=d, generator=g) * 0.1
e_h = torch.empty(E, L, dtype=bf, device=d)
e_y = torch.empty(E, L, dtype=bf, device=d)


stats = torch.empty(E, 4, dtype=torch.float32, device=d)


def fused():
    ops.edge_mlp_rows(base, (gs, idx_s), w2, b2, e_y, row_stats=stats)


def unfused():
    ops.edge_hidden(base, [(gs, idx_s)], e_h, act="swish")
    ops.gemm([(e_h, w2)], e_y, bias=b2, static_weights=True)


for name, fn in (("fused", fused), ("two kernels", unfused)):
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

# ---- clock-stamp timeline of CTA 0 of the fused kernel
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
nt = min(5, int((tr[0][2::3] > 0).sum()))
print("tiles of CTA 0 shown:", nt, "(cycles since the first stamp)")
print("MMA per tile (wait acc_empty start, acc free, all issued):", [(rel(tr[0][3 * i]), rel(tr[0][3 * i + 1]), rel(tr[0][3 * i + 2])) for i in range(nt)])
print("MMA: k-block operands ready at:", [[rel(tr[1][i * KB + k]) for k in range(KB)] for i in range(nt)])
print("producer warp 2 per k-block (wait TMA start, landed, stage signalled):")
for i in range(nt):
    print("   ", [(rel(tr[2][3 * (i * KB + k)]), rel(tr[2][3 * (i * KB + k) + 1]), rel(tr[2][3 * (i * KB + k) + 2])) for k in range(KB)])
print("epilogue warp 10 per tile (wait acc start, acc arrived, -, tile stored):",
      [tuple(rel(tr[3][4 * i + j]) for j in range(4)) for i in range(nt)])
