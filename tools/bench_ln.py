import sys, torch
sys.path.insert(0, ".")
from gencast_flax_nnx_b200 import ops
d = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=d)
def run(name, x, out, so, res=None, n=10):
    for _ in range(3): ops.ln_cond(x, out, so, residual=res)
    torch.cuda.synchronize()
    ts = []
    for i in range(n):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.ln_cond(x, out, so, residual=res); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[len(ts)//2] * 1e-3
    nb = x.numel()*x.element_size() + out.numel()*out.element_size() + (0 if res is None else res.numel()*res.element_size())
    print(f"{name}: {t*1e6:.1f} us, {nb/t/1e9:.0f} GB/s")
L = 512
so = torch.cat([1 + 0.1 * torch.randn(L), torch.randn(L)]).to(d)
x = torch.randn(41472, L, device=d); out = torch.empty(41472, L, dtype=torch.bfloat16, device=d)
run("transformer LN (fp32 -> bf16, 41472 rows)", x, out, so)
y = torch.randn(260640, L, device=d).to(torch.bfloat16); o2 = torch.empty_like(y); r = torch.randn(260640, L, device=d).to(torch.bfloat16)
run("grid LN (bf16 -> bf16, 260640 rows)", y, o2, so)
run("grid LN + residual", y, o2, so, r)
