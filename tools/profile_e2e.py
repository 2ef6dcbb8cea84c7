"""Host-side profile of GenCast.full_sampling (where the e2e overhead over the device time goes).
Usage (GPU box): python tools/profile_e2e.py [config] [members]"""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from gencast_flax_nnx_b200 import configs, gencast
from gencast_flax_nnx_b200.rngs import Rngs
config = sys.argv[1] if len(sys.argv) > 1 else "1deg"
MB = int(sys.argv[2]) if len(sys.argv) > 2 else 4
case = bench.build_case(config, batch=MB)
dev = torch.device("cuda:0")
model = gencast.GenCast(configs.TASK, case["arch"], sampler_config=configs.SamplerConfig(stochastic_churn_rate=0.0),
                        rngs=Rngs(0), params=case["params"], compute_dtype="bf16", device=dev)
for _ in range(2):
    model.full_sampling(case["inputs"], case["targets"], case["forcings"])
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    model.full_sampling(case["inputs"], case["targets"], case["forcings"])
torch.cuda.synchronize()
pr.disable()
print("per call ms:", (time.perf_counter() - t0) / 3 * 1e3)
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
