"""Clock-stamp timeline of one CTA of the tensor-core attention kernel (debug aid).
Usage (on a GPU box): python tools/trace_attention.py [nano|1deg]"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from gencast_flax_nnx_b200 import _lib, configs, graph, ops

name = sys.argv[1] if len(sys.argv) > 1 else "1deg"
res, arch = configs.named_config(name)
st = arch.sparse_transformer_config
mesh = graph.permute_mesh_to_banded(graph.icosphere(arch.mesh_size))
kh = graph.khop_neighbourhoods(mesh, st.attention_k_hop)
order = graph.patch_order(mesh.vertices, 128)
kh = kh.tocsr()[order][:, order].tocsr()
tp, tk, tm = graph.khop_tiles(kh, 128)
V, H, D = kh.shape[0], st.num_heads, st.d_model // st.num_heads
d = torch.device("cuda:0")
qkv = torch.randn(V, 3 * H * D, device=d).to(torch.bfloat16)
out = torch.empty(V, H * D, dtype=torch.bfloat16, device=d)
args = (qkv, out, torch.from_numpy(tp).to(d), torch.from_numpy(tk).to(d), torch.from_numpy(tm.view(np.int32)).to(d), H, D)
lib = _lib.load()
for _ in range(3):
    ops.khop_attention_tiles(*args)
torch.cuda.synchronize()
trace = torch.zeros(8 * 512, dtype=torch.int64, device=d)
lib.gc_debug_set_attention_trace(ctypes.c_void_p(trace.data_ptr()))
ops.khop_attention_tiles(*args)
torch.cuda.synchronize()
lib.gc_debug_set_attention_trace(ctypes.c_void_p(0))
t = trace.cpu().numpy().reshape(8, 512)
T = int(tp[1] - tp[0])
t0 = t[t > 0].min()
rel = lambda x: (x - t0) if x > 0 else -1
print(f"{name}: CTA 0 has T={T} key tiles; times in cycles since first event")
print("TMA issue (after slot free):", [rel(x) for x in t[0][: 2 * T]])
print("MMA S  (K ready, S buffer free):", [(rel(t[1][2 * i]), rel(t[1][2 * i + 1])) for i in range(T)])
print("MMA PV (V ready, P ready):", [(rel(t[2][2 * i]), rel(t[2][2 * i + 1])) for i in range(T)])
print("softmax warp 2 (group 0; wait S start, S arrived, P buffer free) per even tile:",
      [(rel(t[3][2 * i]), rel(t[3][2 * i + 1]), rel(t[4][i])) for i in range(0, T, 2)])
print("epilogue wait O (start, arrived):", rel(t[5][0]), rel(t[5][1]))
