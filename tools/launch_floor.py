"""Per-kernel floor inside a CUDA graph: chains of dependent tiny launches (debug aid)."""
import sys, torch
sys.path.insert(0, ".")
from gencast_flax_nnx_b200 import ops
d = torch.device("cuda:0")
def chain(fn, n=200):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / n
for (M, N, K) in [(128, 128, 64), (2562, 256, 256), (2562, 768, 256), (2562, 2048, 256), (2562, 256, 2048), (10512, 256, 256)]:
    a = torch.randn(M, K, device=d).to(torch.bfloat16); w = torch.randn(N, K, device=d).to(torch.bfloat16)
    o1 = torch.empty(M, N, dtype=torch.bfloat16, device=d)
    print(f"gemm {M}x{N}x{K}: {chain(lambda: ops.gemm([(a, w)], o1)):.2f} us per launch")
x = torch.randn(2562, 256, device=d); h = torch.empty(2562, 256, dtype=torch.bfloat16, device=d); so = torch.randn(512, device=d)
print(f"ln_cond 2562x256: {chain(lambda: ops.ln_cond(x, h, so)):.2f} us per launch")
x2 = torch.randn(128, 256, device=d); h2 = torch.empty(128, 256, dtype=torch.bfloat16, device=d)
print(f"ln_cond 128x256: {chain(lambda: ops.ln_cond(x2, h2, so)):.2f} us per launch")
