"""Experiment: 4 members as one batched graph vs two concurrent graphs of 2 members on two streams."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from gencast_flax_nnx_b200.engine import DenoiserEngine, SamplerEngine, noise_schedule
dev = torch.device("cuda:0")
case = bench.build_case("1deg", batch=4)
graphs = bench.oracle_graph(case)
sig = noise_schedule(80.0, 0.03, 20, 7.0)
mm = lambda a: np.ascontiguousarray(np.transpose(a, (1, 0, 2))).reshape(-1, a.shape[-1])
def make(members, sl):
    eng = DenoiserEngine(graphs, case["arch"], case["params"], case["layout"], compute_dtype="bf16", device=dev, members=members)
    se = SamplerEngine(eng, sig)
    eng.set_constant_features(mm(case["inp_nodes"][:, sl]), mm(case["frc_nodes"][:, sl]))
    return eng, se
def timeit(fn, n=4):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
e4, s4 = make(4, slice(0, 4))
n4 = torch.randn(e4.Gt, e4.n_out, device=dev)
t4 = timeit(lambda: s4.sample(n4))
print(f"one graph, 4 members: {t4:.1f} ms per step -> {4e3 / t4:.3f} member-steps/s")
del e4, s4; torch.cuda.empty_cache()
parts = [make(2, slice(0, 2)), make(2, slice(2, 4))]
noise = [torch.randn(p[0].Gt, p[0].n_out, device=dev) for p in parts]
streams = [torch.cuda.Stream(device=dev) for _ in parts]
for (e, s), nz in zip(parts, noise): s.sample(nz)          # capture
torch.cuda.synchronize()
def both():
    main = torch.cuda.current_stream()
    for st, (e, s), nz in zip(streams, parts, noise):
        st.wait_stream(main)
        with torch.cuda.stream(st):
            s._graph.replay()
    for st in streams: main.wait_stream(st)
t2 = timeit(both)
print(f"two graphs of 2 members on two streams: {t2:.1f} ms per step -> {4e3 / t2:.3f} member-steps/s")
