"""Clock-stamp timeline of CTA 0 (rank 0 of pair 0) of the column-split CTA-pair variant of gc_edge_mlp_sum3
(GENCAST_EDGE_PAIR=1): MMA issuer, producers, forwarder, epilogue.  Usage (GPU box, repo root): python tools/trace_edge_pair.py"""
import os, sys
os.environ["GENCAST_EDGE_PAIR"] = "1"
sys.argv = [sys.argv[0], "4"]
exec(open("tools/bench_edge_fused.py").read().split("# ---- clock-stamp timeline")[0])
import ctypes
from gencast_flax_nnx_b200 import _lib
lib = _lib.load()
trace = torch.zeros(8 * 512, dtype=torch.int64, device=d)
lib.gc_debug_set_edge_fused_trace(ctypes.c_void_p(trace.data_ptr()))
fused(); torch.cuda.synchronize()
lib.gc_debug_set_edge_fused_trace(ctypes.c_void_p(0))
tr = trace.cpu().numpy().reshape(8, 512)
t0 = tr[tr > 0].min(); rel = lambda x: int(x - t0) if x > 0 else -1
KB = 8; nt = 5
print("MMA per tile (wait acc_empty start, acc free, all issued):", [(rel(tr[0][3*i]), rel(tr[0][3*i+1]), rel(tr[0][3*i+2])) for i in range(nt)])
print("MMA k-block ready:", [[rel(tr[1][i*KB+k]) for k in range(KB)] for i in range(nt)])
print("producer (wait TMA start, landed, signalled) per own k-block:")
for i in range(nt): print("   ", [(rel(tr[2][3*(i*4+n)]), rel(tr[2][3*(i*4+n)+1]), rel(tr[2][3*(i*4+n)+2])) for n in range(4)])
print("forwarder (local ready, peer go-ahead):", [[(rel(tr[5][2*(i*4+n)]), rel(tr[5][2*(i*4+n)+1])) for n in range(4)] for i in range(nt)])
print("epilogue (wait acc start, arrived, stats exchanged, stored):", [tuple(rel(tr[3][4*i+j]) for j in range(4)) for i in range(nt)])
