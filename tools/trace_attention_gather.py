"""Clock-stamp timeline of CTA 0 of the gather attention kernel (debug aid).
Usage (on a GPU box): python tools/trace_attention_gather.py [nano|1deg] [members]"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gencast_flax_nnx_b200 import _lib, configs, graph, ops

name = sys.argv[1] if len(sys.argv) > 1 else "1deg"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
res, arch = configs.named_config(name)
st = arch.sparse_transformer_config
mesh = graph.permute_mesh_to_banded(graph.icosphere(arch.mesh_size))
kh = graph.khop_neighbourhoods(mesh, st.attention_k_hop)
order = graph.patch_order(mesh.vertices, 128)
kh = kh.tocsr()[order][:, order].tocsr()
sp, keys, cm, work = graph.khop_compact_steps(kh, 128, 64)
V, H, D = kh.shape[0], st.num_heads, st.d_model // st.num_heads
Vp = -(-V // 128) * 128
ns = int(sp[-1])
sp = np.concatenate([sp[:-1].astype(np.int64) + b * ns for b in range(B)] + [[B * ns]]).astype(np.int32)
keys = (keys[None, :].astype(np.int64) + (np.arange(B) * Vp)[:, None]).reshape(-1).astype(np.int32)
import os
order = os.environ.get("GENCAST_ATT_ORDER", "natural")
work = {"sorted": np.argsort(-np.diff(sp), kind="stable"), "natural": np.arange(len(sp) - 1),
        "reverse": np.arange(len(sp) - 1)[::-1]}[order].astype(np.int32).copy()
d = torch.device("cuda:0")
qkv = torch.randn(B * Vp, 3 * H * D, device=d).to(torch.bfloat16)
out = torch.empty(B * Vp, H * D, dtype=torch.bfloat16, device=d)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(d)
args = (qkv, out, t(sp), t(keys), t(cm.view(np.int32).reshape(-1)), t(work), H, D, ns)
lib = _lib.load()
for _ in range(3):
    ops.khop_attention_gather(*args)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    ops.khop_attention_gather(*args)
b.record()
torch.cuda.synchronize()
print(f"{name} x{B}: {a.elapsed_time(b) / 20 * 1e3:.1f} us per launch, {len(work) * H} CTAs, {int(sp[-1])} steps")
trace = torch.zeros(8 * 512, dtype=torch.int64, device=d)
lib.gc_debug_set_attention_gather_trace(ctypes.c_void_p(trace.data_ptr()))
ops.khop_attention_gather(*args)
torch.cuda.synchronize()
lib.gc_debug_set_attention_gather_trace(ctypes.c_void_p(0))
tr = trace.cpu().numpy().reshape(8, 512)
qt = int(work[0])
T = int(sp[qt + 1] - sp[qt])
t0 = tr[6][0]
rel = lambda x: int(x - t0) if x > 0 else -1
print(f"CTA 0: query tile {qt}, T={T} steps; cycles since kernel entry")
print("loader warp 5: slot free -> copies issued at:", [rel(x) for x in tr[0][: 2 * T]])

print("MMA S  (wait K start, K ready):", [(rel(tr[1][2 * i]), rel(tr[1][2 * i + 1])) for i in range(T)])
print("MMA PV (start, V ready, P ready):", [(rel(tr[2][3 * i]), rel(tr[2][3 * i + 1]), rel(tr[2][3 * i + 2])) for i in range(T)])
print("softmax warp 1 (wait S start, S arrived, P signalled):", [(rel(tr[3][3 * i]), rel(tr[3][3 * i + 1]), rel(tr[3][3 * i + 2])) for i in range(T)])
print("epilogue (wait O start, O ready, stores done, CTA end):", [rel(x) for x in tr[4][:4]])
