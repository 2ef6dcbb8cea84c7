"""Times gc_ln_cond_segment_sum at the GenCast 1 deg x 4 members shapes (synthetic degrees of the same statistics).
Usage (GPU box): python tools/bench_segsum.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from gencast_flax_nnx_b200 import ops
from gencast_flax_nnx_b200.graph import csr_by_receiver

d = torch.device("cuda:0")
L = 512
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)


def run(name, y, out, so, rp, perm, stats=None):
    for _ in range(3):
        ops.ln_cond_segment_sum(y, out, so, rp, perm, row_stats=stats)
    torch.cuda.synchronize()
    ts = []
    for i in range(10):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.ln_cond_segment_sum(y, out, so, rp, perm, row_stats=stats); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = float(np.median(ts)) * 1e-3
    nbytes = y.numel() * y.element_size() + out.numel() * out.element_size() + 4 * (rp.numel() + (perm.numel() if perm is not None else 0))
    print(f"{name}: {t*1e6:.1f} us, {nbytes/t/1e9:.0f} GB/s ({nbytes/1e6:.0f} MB)")


rng = np.random.default_rng(0)
so = torch.cat([1 + 0.1 * torch.randn(L), torch.randn(L)]).to(d)
# mesh2grid: 3 edges per receiver, already sorted
G = 65160 * 4
y = torch.randn(3 * G, L, device=d).to(torch.bfloat16)
out = torch.empty(G, L, dtype=torch.bfloat16, device=d)
rp = torch.arange(0, 3 * G + 1, 3, dtype=torch.int32, device=d)
run("m2g (deg 3, sorted)", y, out, so, rp, None)
# grid2mesh: mean degree ~10, a few polar receivers with ~600, edges listed through a permutation
V, E = 10242, 101892
deg = rng.poisson(8.8, V) + 1
deg[:12] = 594
deg = (deg * (E / deg.sum())).astype(np.int64)
deg[rng.choice(V, E - deg.sum(), replace=False)] += 1      # spread the rounding remainder
recv1 = rng.permutation(np.repeat(np.arange(V), deg))
recv = np.concatenate([recv1 + b * V for b in range(4)])
rp_np, perm_np = csr_by_receiver(recv, 4 * V)
y = torch.randn(len(recv), L, device=d).to(torch.bfloat16)
out = torch.empty(4 * V, L, dtype=torch.bfloat16, device=d)
run("g2m (perm)", y, out, so, torch.from_numpy(rp_np).to(d), torch.from_numpy(perm_np).to(d))
# same edges physically receiver-sorted
ys = y[torch.from_numpy(perm_np).long().to(d)].contiguous()
run("g2m (sorted rows)", ys, out, so, torch.from_numpy(rp_np).to(d), None)
# the same with the rows' statistics supplied by the producer (gc_edge_mlp_rows writes them; synthetic values here)
yf = ys.float()
h = L // 2
stats = torch.stack([yf[:, :h].sum(1), (yf[:, :h] ** 2).sum(1), yf[:, h:].sum(1), (yf[:, h:] ** 2).sum(1)], dim=1).contiguous()
run("g2m (sorted rows, statistics supplied)", ys, out, so, torch.from_numpy(rp_np).to(d), None, stats)
