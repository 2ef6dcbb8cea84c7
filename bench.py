#!/usr/bin/env python
"""Benchmark of the GenCast sampling hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config 1deg|nano|0p25deg|tiny] [--members-per-gpu B]
  python bench.py --impl reference ...      # the reference algorithm's CPU restatement (oracle/) on host cores

One "step" = one 12 h forecast step of the ensemble members resident on a GPU: the 20-level
DPM-Solver++ 2S loop, 40 denoiser evaluations per member (the reference evaluates and discards
the 40th; so do we).  Default workload = the configuration the north-star target is quoted on,
GenCast 1 deg with BASELINE.json configs[3]'s sharding (32 members over 8 GPUs = 4 members per
GPU, evaluated together: member-major row blocks through the same kernels); members are
independent, so N GPUs run 4 N members (weak scaling, no collective on the data path; with
N > 1 the per-step ensemble sum / sum-of-squares is all-reduced over NCCL).  The same line also
carries configs[1] (nano-GenCast 2.5 deg, one member per GPU) under "also".
`value` = members x steps / max-over-ranks device time with inputs resident in HBM.
`e2e` = the same through GenCast.full_sampling with host Datasets (H2D of the step's
inputs from pinned memory and D2H of the prediction inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "12h_forecast_steps_per_sec"
UNIT = "member-steps/s"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=p["hbm_gbs"], tensor_burst=p["bf16_tflops"], tensor_sustained=p["bf16_tflops_sustained"],
                    source="measured")
    except Exception:
        return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_case(config: str, seed: int = 0, batch: int = 1):
    from gencast_flax_nnx_b200 import configs, graph, params, stacking, synthetic
    from gencast_flax_nnx_b200.engine import ChannelLayout
    res, arch = configs.named_config(config)
    lat, lon = graph.regular_grid(res)
    inputs, targets, forcings = synthetic.make_example(lat, lon, batch=batch, seed=seed)
    sizes = dict(targets.sizes)
    inp_nodes, _ = stacking.dataset_to_nodes(inputs, sizes)
    frc_nodes, frc_layout = stacking.dataset_to_nodes(forcings, sizes)
    layout = ChannelLayout(num_input_channels=inp_nodes.shape[-1], forcing_vars=tuple(frc_layout),
                           target_vars=tuple(stacking.channel_layout(targets)))
    shapes = params.param_shapes(arch, layout.num_data_channels, layout.num_targets)
    p = params.init_perturbed(shapes, seed=1)     # random O(1/sqrt(fan_in)) weights
    return dict(res=res, arch=arch, lat=lat, lon=lon, inputs=inputs, targets=targets, forcings=forcings,
                inp_nodes=inp_nodes, frc_nodes=frc_nodes, frc_layout=frc_layout, layout=layout, params=p)


def workload_name(config: str) -> str:
    return {"nano": "nano-GenCast 2.5deg (73x144 grid, mesh 4, L=256, 16 layers, k-hop 8), "
                    "12 h step = 20-level DPM-Solver++ 2S",
            "1deg": "GenCast 1deg (181x360 grid, mesh 5, L=512, 16 layers, k-hop 8), "
                    "12 h step = 20-level DPM-Solver++ 2S",
            "0p25deg": "GenCast 0.25deg (721x1440 grid, mesh 6, L=512, 16 layers, k-hop 8), "
                       "12 h step = 20-level DPM-Solver++ 2S",
            "tiny": "test-size GenCast 10deg (19x36 grid, mesh 2, L=128, 2 layers), 12 h step = 20-level DPM-Solver++ 2S",
            }[config]


DEFAULT_MEMBERS_PER_GPU = {"1deg": 4}      # BASELINE.json configs[3]: 32 members / 8 GPUs


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm as written, restated in torch fp32 (oracle/), on host cores
# ----------------------------------------------------------------------------------------------

def cpu_solver_iteration_seconds(case, repeats: int, warmup: int = 0):
    """Times one solver iteration (2 denoiser evaluations + updates) of the oracle on all host cores."""
    import torch
    from oracle import gencast_oracle as o
    torch.set_num_threads(os.cpu_count() or 1)
    g = case["graphs_oracle"]
    dt = torch.float32
    G, n_out = case["inp_nodes"].shape[0], case["layout"].num_targets
    rng = np.random.default_rng(7)
    tv = sorted(case["layout"].target_vars)
    x, i = {}, 0
    noise = torch.as_tensor(rng.standard_normal((G, 1, n_out)).astype(np.float32)) * 80.0
    for n, c in tv:
        x[n] = noise[:, :, i:i + c]
        i += c
    frc, i = {}, 0
    for n, c in case["frc_layout"]:
        frc[n] = torch.as_tensor(np.ascontiguousarray(case["frc_nodes"][:, :1, i:i + c]))      # one member
        i += c
    inp = torch.as_tensor(np.ascontiguousarray(case["inp_nodes"][:, :1]))
    st = case["arch"].sparse_transformer_config
    arch = dict(num_layers=st.num_layers, num_heads=st.num_heads)
    sig = [80.0, 60.0]
    times = []
    with torch.no_grad():
        for r in range(warmup + repeats):
            t0 = time.perf_counter()
            o.dpm_solver_2s(case["params"], g, arch, inp, frc, x, sig, dt, num_steps=1)
            if r >= warmup:
                times.append(time.perf_counter() - t0)
    return times


def oracle_graph(case):
    from gencast_flax_nnx_b200 import graph
    st = case["arch"].sparse_transformer_config
    g = graph.build_denoiser_graphs(case["lat"], case["lon"], case["arch"].mesh_size, st.attention_k_hop)
    case["graphs"] = g
    case["graphs_oracle"] = dict(g2m_grid_feat=g.g2m_grid_feat, g2m_mesh_feat=g.g2m_mesh_feat,
                                 g2m_edge_feat=g.g2m_edge_feat, g2m_senders=g.g2m_senders,
                                 g2m_receivers=g.g2m_receivers, m2g_senders=g.m2g_senders,
                                 m2g_receivers=g.m2g_receivers, m2g_edge_feat=g.m2g_edge_feat, khop=g.khop)
    return g


def cpu_full_step_seconds(case, repeats: int = 1):
    """Times whole 12 h steps (all 20 solver iterations, 39 denoiser evaluations + updates; the reference's 40th,
    discarded evaluation is added as one more evaluation's time) of the oracle on all host cores: no extrapolation."""
    import torch
    from oracle import gencast_oracle as o
    torch.set_num_threads(os.cpu_count() or 1)
    g = case["graphs_oracle"]
    dt = torch.float32
    G, n_out = case["inp_nodes"].shape[0], case["layout"].num_targets
    rng = np.random.default_rng(7)
    x, i = {}, 0
    sig = o.noise_schedule(80.0, 0.03, 20, 7.0)
    noise = torch.as_tensor(rng.standard_normal((G, 1, n_out)).astype(np.float32)) * float(sig[0])
    for n, c in sorted(case["layout"].target_vars):
        x[n] = noise[:, :, i:i + c]
        i += c
    frc, i = {}, 0
    for n, c in case["frc_layout"]:
        frc[n] = torch.as_tensor(np.ascontiguousarray(case["frc_nodes"][:, :1, i:i + c]))
        i += c
    inp = torch.as_tensor(np.ascontiguousarray(case["inp_nodes"][:, :1]))
    st = case["arch"].sparse_transformer_config
    arch = dict(num_layers=st.num_layers, num_heads=st.num_heads)
    times = []
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            o.dpm_solver_2s(case["params"], g, arch, inp, frc, x, sig, dt)
            times.append((time.perf_counter() - t0) * 40.0 / 39.0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    case = build_case(args.config)
    oracle_graph(case)
    iters_per_step = 20
    also = None
    if args.config in ("nano", "tiny"):
        # the reference's own CPU-runnable case (BASELINE configs[0]): whole steps, nothing extrapolated
        samples = max(1, min(args.steps, 2))
        times = cpu_full_step_seconds(case, repeats=samples)
        t_step = float(np.mean(times))
        extrapolated = False
        sample = (f"{samples} whole 12 h member-step(s) (20 solver iterations, 40 denoiser evaluations) of one member, "
                  "torch-fp32 restatement of the reference algorithm on all host threads; not extrapolated")
    else:
        # bounded: at 1 deg one solver iteration takes ~20 s on 16 host cores, so at most 3 are timed (after 1 warm-up)
        samples = max(1, min(args.steps, 3))
        times = cpu_solver_iteration_seconds(case, repeats=samples, warmup=min(args.warmup, 1))
        t_step = float(np.mean(times)) * iters_per_step
        extrapolated = True
        sample = (f"{samples} timed samples, each 1 of the 20 solver iterations (2 denoiser evaluations + updates) of one "
                  "member, torch-fp32 restatement of the reference algorithm (dense tri-block attention, [e|s|r] concat, "
                  "scatter-add) on all host threads; 12 h member-step time = 20 x the mean sample (extrapolated)")
        if args.config == "1deg" and not args.no_secondary:
            nano = build_case("nano")
            oracle_graph(nano)
            tn = float(np.mean(cpu_full_step_seconds(nano, repeats=1)))
            also = {"nano_2p5deg_one_member": {"workload": workload_name("nano"), "value": 1.0 / tn, "unit": UNIT,
                                               "ms_per_step": tn * 1e3, "extrapolated": False,
                                               "sample": "one whole 12 h member-step (40 denoiser evaluations), timed in full"}}
    value = 1.0 / t_step
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "extrapolated": extrapolated,
            "config": {"workload": workload_name(args.config), "members": 1,
                       "note": "CPU arm: one member at a time; member-steps/s does not depend on how members are grouped"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample,
                             "extrapolated": extrapolated},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if also is not None:
        line["also"] = also
    emit(line)


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------

def measure(args, config: str, MB: int, steps: int, warmup: int, dev, rank: int, world: int, detailed: bool):
    """Device-timed steps, e2e steps and (detailed) the per-kernel roofline leg of one workload."""
    import torch
    import torch.distributed as dist
    from gencast_flax_nnx_b200 import configs, gencast, ops
    from gencast_flax_nnx_b200.engine import DenoiserEngine, SamplerEngine, noise_schedule
    from gencast_flax_nnx_b200.parallel import EnsembleStatistics
    from gencast_flax_nnx_b200.rngs import Rngs

    case = build_case(config, batch=MB)
    graphs = oracle_graph(case)
    eng = DenoiserEngine(graphs, case["arch"], case["params"], case["layout"], compute_dtype=args.dtype, device=dev,
                         members=MB)
    sigmas = noise_schedule(80.0, 0.03, 20, 7.0)
    se = SamplerEngine(eng, sigmas, evaluate_discarded_call=True)
    member_major = lambda a: np.ascontiguousarray(np.transpose(a, (1, 0, 2))).reshape(-1, a.shape[-1])
    eng.set_constant_features(member_major(case["inp_nodes"]), member_major(case["frc_nodes"]))
    G, C = eng.Gt, eng.n_out
    gen = torch.Generator(device=dev); gen.manual_seed(1234 + rank)
    noises = [torch.randn(G, C, generator=gen, device=dev) for _ in range(max(steps, 1))]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    Gm = eng.G                                   # grid nodes of one member
    stats = EnsembleStatistics((Gm, C), dev)

    def one_step(noise):
        out = se.sample(noise, use_graph=True)
        if world > 1:
            # ensemble mean / spread over the members of all ranks: the local members' sum and sum of squares
            # (gc_ensemble_accumulate per member) + ONE NCCL all-reduce of [2, G, 82] floats
            stats.reset()
            for b in range(MB):
                stats.add(out[b * Gm:(b + 1) * Gm])
            stats.finalize(total_members=world * MB)
        return out

    for i in range(warmup):
        one_step(noises[i % len(noises)])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
    with ClockSampler(dev.index) as clocks:
        for i in range(steps):
            flush.fill_(i & 0xFF)           # evict L2 between timed steps (not timed)
            starts[i].record()
            one_step(noises[i])
            ends[i].record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    dev_ms = sum(a.elapsed_time(b) for a, b in zip(starts, ends))
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    extra = {}
    if world > 1 and detailed:
        # correctness of the reduced statistics on real GPUs: the all-reduced mean / spread of a slice of grid nodes
        # against an all-gather of every member's slice; and the fair CRPS over NCCL, timed once
        from gencast_flax_nnx_b200.parallel import fair_crps
        out = one_step(noises[0])
        mean, spread, m = stats.finalize(total_members=world * MB)
        sl = out.reshape(MB, Gm, C)[:, :256].contiguous()
        parts = [torch.empty_like(sl) for _ in range(world)]
        dist.all_gather(parts, sl)
        allm = torch.cat(parts, 0).double()
        d_mean = float((allm.mean(0) - mean[:256].double()).abs().max())
        d_spread = float((allm.std(0, unbiased=True) - spread[:256].double()).abs().max())
        truth = torch.zeros(Gm, C, device=dev)
        lat_w = torch.cos(torch.deg2rad(torch.as_tensor(np.repeat(case["lat"], len(case["lon"])), device=dev))).clamp_min(0).float()
        fair_crps(out.reshape(MB, Gm, C), truth, lat_w)
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        crps = fair_crps(out.reshape(MB, Gm, C), truth, lat_w)
        b.record()
        torch.cuda.synchronize()
        tc = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        extra["stats_check"] = {"members": m, "max_abs_diff_mean": d_mean, "max_abs_diff_spread": d_spread,
                                "ok": bool(d_mean < 1e-4 and d_spread < 1e-3),
                                "how": "all-reduced mean / spread of 256 grid nodes vs an all-gather of every member"}
        extra["crps"] = {"ms": float(tc.item()), "members": world * MB, "finite": bool(torch.isfinite(crps).all()),
                         "how": "fair CRPS of all members over NCCL (all-gather of members, gc_fair_crps on each rank's "
                                "grid slice, all-reduce of the weighted sums), once, max over ranks"}
    res = {"value": world * MB * steps / (total_ms / 1e3), "ms_per_step": total_ms / steps, "clocks": clocks.summary(), **extra,
           "members_per_gpu": MB, "evals": se.num_network_evaluations, "launches_per_step": se.launches_per_step,
           "denoiser_fwd_ms": total_ms / steps / se.num_network_evaluations}

    # ---- e2e through the public API: host Datasets in, host Dataset out
    sc = configs.SamplerConfig(stochastic_churn_rate=0.0)
    model = gencast.GenCast(configs.TASK, case["arch"], sampler_config=sc, rngs=Rngs(rank), params=case["params"],
                            compute_dtype=args.dtype, device=dev)
    model.denoiser._engine = eng                                   # share the resident weights / graph tables
    model.denoiser._grid_key = (np.asarray(case["inputs"].coords["lat"]).tobytes(),
                                np.asarray(case["inputs"].coords["lon"]).tobytes())
    model._sampler._engine = se
    e2e_steps = max(2, min(steps, 5))
    from gencast_flax_nnx_b200.device_stacking import pin_dataset
    inputs_h, forcings_h = pin_dataset(case["inputs"]), pin_dataset(case["forcings"])     # page-locked host buffers
    for _ in range(2):
        model.full_sampling(inputs_h, case["targets"], forcings_h)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pred = model.full_sampling(inputs_h, case["targets"], forcings_h)
        checksum = float(pred["2m_temperature"].data[0, 0, 0, 0]) if "2m_temperature" in pred else 0.0   # result is on the host
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["e2e"] = {"value": world * MB * e2e_steps / float(t.item()), "unit": UNIT,
                  "h2d_bytes_per_step": sum(int(np.prod(v.shape)) for ds in (case["inputs"], case["forcings"])
                                            for v in ds.data_vars.values()) * 4,
                  "d2h_bytes_per_step": C * G * 4, "steps": e2e_steps,
                  "api": "GenCast.full_sampling(inputs, targets_template, forcings)",
                  "host_buffers": "inputs / forcings Datasets in page-locked host memory (H2D every step); the "
                                  "prediction Dataset is host arrays (one D2H per step into page-locked memory)"}
    res["case"] = case
    if args.rollout_steps > 0 and detailed:
        # autoregressive rollout with the input window resident on the GPU (rollout.device_chunked_prediction):
        # inputs uploaded once, forcings per step, predictions downloaded per step
        from gencast_flax_nnx_b200 import rollout, synthetic
        R = args.rollout_steps
        _, tgt_r, frc_r = synthetic.make_example(case["lat"], case["lon"], batch=MB, seed=0, num_target_steps=R)
        frc_r = pin_dataset(frc_r)
        for _ in rollout.device_chunked_prediction_generator(model, inputs_h, tgt_r.isel(time=slice(0, 1)), frc_r.isel(time=slice(0, 1))):
            pass
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for pred in rollout.device_chunked_prediction_generator(model, inputs_h, tgt_r, frc_r):
            pass
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["rollout"] = {"steps": R, "value": world * MB * R / float(t.item()), "unit": UNIT,
                          "api": "rollout.device_chunked_prediction_generator (window on the GPU; forcings H2D and "
                                 "prediction D2H every step)"}
    if not detailed or rank != 0:
        del model, se, eng, stats, flush, noises
        torch.cuda.empty_cache()
        return res

    # ---- roofline leg (rank 0): CUDA-event pair around every launch of denoiser evaluations of the
    # sampling step, replayed eagerly.  The launches are enqueued behind a device-side delay so that
    # the kernels run back to back from the queue: the events then measure device durations, not the
    # host's launch rate (which at the small grid is slower than the kernels themselves).
    peaks = _peaks()
    rec = ops.Recorder()
    se.sample(noises[0], use_graph=False)
    torch.cuda.synchronize()
    for j in (3, 17, 31):                                   # three noise levels of the schedule
        torch.cuda._sleep(int(2.5e7))                       # ~13 ms of head start for the host
        ops.set_recorder(rec)
        f = eng.forward(se.ctx[j])
        ops.dpm_update(f, se.x, se.x, se.sched[j], se.x_mid, eng.xin, C)
        ops.set_recorder(None)
        torch.cuda.synchronize()
    agg = rec.summary()
    tot_ms = sum(d["ms"] for d in agg.values())
    kernels = {}
    for name, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        avg_ms = d["ms"] / d["launches"]
        k = {"launches": d["launches"], "avg_us": 1e3 * avg_ms, "share": d["ms"] / tot_ms}
        if d["flops"] > 0:
            k["tflops"] = d["flops"] / d["ms"] / 1e9
        k["gbs"] = d["bytes"] / d["ms"] / 1e6
        if d["flops"] == 0:
            k["hbm_frac"] = k["gbs"] / peaks["hbm"]     # HBM-bound kernels: algorithmic bytes / time over the measured copy peak
        kernels[name] = k
    ka = kernels.get("khop_attention_gather")
    if ka is not None and getattr(eng, "attention_kind", "") == "gather":
        # `tflops` counts the exact k-hop pattern (4 nnz(adj^k) L); what the tensor cores execute is dense 128 x 64 tiles
        # over the compacted key steps: S = Q K^T and O += P V, 2 x (2 * 128 * 64 * head_dim) FLOP per step and head
        exec_flops = float(eng.num_attention_steps) * eng.H * 4.0 * 128 * 64 * eng.head_dim
        ka["tflops_executed"] = exec_flops / (ka["avg_us"] * 1e-6) / 1e12
        ka["executed_frac_of_tensor_peak"] = ka["tflops_executed"] / peaks["tensor_sustained"]
    dom = next(iter(kernels))
    domk = kernels[dom]
    tensor_bound = dom.startswith("gemm_bf16") or dom.startswith("khop_attention_tc")
    if tensor_bound:
        roof = {"kernel": dom, "bound": "tensor", "achieved": domk["tflops"], "peak": peaks["tensor_sustained"],
                "unit": "TFLOP/s", "frac": domk["tflops"] / peaks["tensor_sustained"], "traffic": None}
    else:
        roof = {"kernel": dom, "bound": "hbm", "achieved": domk["gbs"], "peak": peaks["hbm"], "unit": "GB/s",
                "frac": domk["gbs"] / peaks["hbm"], "traffic": None}
    try:
        # dram__bytes_read.sum + dram__bytes_write.sum per launch of every kernel family from an ncu capture of one
        # evaluation of the same workload (profiles/r02_dram_traffic.json names the command); not measured live
        with open(os.path.join(ROOT, "profiles", "r02_dram_traffic.json")) as f:
            tr = json.load(f)
        if tr["workload"] == config and tr["members_per_gpu"] == MB:
            for name, k in kernels.items():
                fam = tr["families"].get(name)
                if fam is not None:
                    k["dram_bytes_per_launch"] = fam["dram_bytes_per_launch"]
            fam = tr["families"].get(dom)
            if fam is not None:
                roof["traffic"] = fam["dram_bytes_per_launch"]
                roof["traffic_unit"] = "bytes per launch (ncu capture, profiles/r02_dram_traffic.json)"
                roof["algorithmic_bytes_per_launch"] = domk["gbs"] * 1e9 * domk["avg_us"] * 1e-6
    except Exception:
        pass
    roof["peak_source"] = peaks["source"] + (" (sustained bf16 GEMM: the kernel is timed inside a long step)"
                                             if tensor_bound else " (copy bandwidth)")
    roof["avg_launch_us"] = domk["avg_us"]
    roof["share_of_step"] = domk["share"]
    # whole-evaluation figure: algorithmic FLOPs of one member's forward (exact k-hop attention)
    f_alg = algorithmic_flops(eng)
    roof["forward_alg_tflops_per_member"] = f_alg / 1e12
    roof["forward_achieved_tflops"] = MB * f_alg / (res["denoiser_fwd_ms"] * 1e-3) / 1e12
    roof["forward_frac_of_peak"] = roof["forward_achieved_tflops"] / peaks["tensor_sustained"]
    # F_exec: the FLOPs the kernels actually issue per evaluation (tensor-core GEMMs + exact k-hop attention).  It is
    # smaller than F_alg because everything that depends on the weights and the noise level only (edge / mesh-node
    # embedders, the edge part of the edge MLPs' first layer) is tabulated per level instead of recomputed.
    f_exec = sum(d["flops"] for d in agg.values()) / 3.0
    roof["forward_exec_tflops_per_member"] = f_exec / MB / 1e12
    roof["forward_exec_tflops"] = f_exec / (res["denoiser_fwd_ms"] * 1e-3) / 1e12
    roof["forward_exec_frac_of_peak"] = roof["forward_exec_tflops"] / peaks["tensor_sustained"]
    # the same GEMM shapes through cuBLAS (torch.matmul), back to back under the same power cap: how much of the gap
    # to the 8192^3 peak is the shape (M x 512 x 512 problems are L2 / HBM fed) and how much the kernel
    try:
        roof["gemm_vs_cublas_same_shape"] = gemm_vs_cublas(eng, dev)
    except Exception as e:                                   # a comparison, never a reason to lose the line
        roof["gemm_vs_cublas_same_shape"] = {"error": str(e)[:200]}
    res["roofline"], res["kernels"] = roofline_clean(roof), kernels

    # ---- single denoiser evaluation latency (graph-free, events)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx = se.ctx[5]
    for _ in range(3):
        eng.forward(ctx)
    a.record()
    for _ in range(10):
        eng.forward(ctx)
    b.record()
    torch.cuda.synchronize()
    res["denoiser_fwd_ms_eager_launch"] = a.elapsed_time(b) / 10
    del model, se, eng, stats, flush, noises
    torch.cuda.empty_cache()
    return res


def gemm_vs_cublas(eng, dev, seconds: float = 0.25):
    """Sustained TFLOP/s of gc_gemm and of torch.matmul (cuBLAS) on the transformer / grid-MLP GEMM shapes of this
    workload; cuBLAS is the comparison here, never on the product path."""
    import torch
    from gencast_flax_nnx_b200 import ops
    L, F, V, G = eng.L, eng.F, eng.Vt, eng.Gt
    out = {}
    for name, m, n, k in (("qkv", V, 3 * L, L), ("out_proj", V, L, L), ("ffw_in", V, F, L), ("ffw_out", V, L, F), ("grid_mlp", G, L, L)):
        a = torch.randn(m, k, device=dev).to(torch.bfloat16)
        w = (torch.randn(n, k, device=dev) / 22.6).to(torch.bfloat16)
        o = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        tf = {}
        for impl, fn in (("gc_gemm", lambda: ops.gemm([(a, w)], o, static_weights=True)), ("cublas", lambda: torch.matmul(a, w.t(), out=o))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters, start = 0, time.perf_counter()
            t0.record()
            while time.perf_counter() - start < seconds:
                for _ in range(20):
                    fn()
                iters += 20
                torch.cuda.synchronize()
            t1.record()
            torch.cuda.synchronize()
            tf[impl] = 2.0 * m * n * k * iters / (t0.elapsed_time(t1) * 1e-3) / 1e12
        out[name] = {"m": m, "n": n, "k": k, "gc_gemm_tflops": round(tf["gc_gemm"], 1), "cublas_tflops": round(tf["cublas"], 1),
                     "ratio": round(tf["gc_gemm"] / tf["cublas"], 3)}
    return out


def algorithmic_flops(eng) -> float:
    """F_alg of one member's denoiser evaluation: every MLP as 2 n (i h + h o) with the reference's
    operand widths ([e|s|r] = 3L, [n|agg] = 2L), attention at the exact k-hop nnz; the dead mesh-node MLP excluded."""
    L, G, V, E1, E2, F, NL = eng.L, eng.G, eng.V, eng.E1, eng.E2, eng.F, eng.NL
    cin = 3 + eng.layout.num_data_channels
    mlp = lambda n, i, h, o: 2.0 * n * (i * h + h * o)
    enc = mlp(G, cin, L, L) + mlp(V, cin, L, L) + mlp(E1, 4, L, L) + mlp(E1, 3 * L, L, L) + mlp(V, 2 * L, L, L) + mlp(G, L, L, L)
    dec = mlp(E2, 4, L, L) + mlp(E2, 3 * L, L, L) + mlp(G, 2 * L, L, L) + mlp(G, L, L, eng.n_out)
    nnz = eng.khop_nnz / eng.B
    proc = NL * (8.0 * V * L * L + 4.0 * V * L * F + 4.0 * nnz * L)
    return enc + proc + dec


def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")       # keep stdout to the single JSON line
        dist.init_process_group("nccl", device_id=dev)

    MB = args.members_per_gpu
    main = measure(args, args.config, MB, args.steps, args.warmup, dev, rank, world, detailed=True)
    also = None
    if args.config == "1deg" and not args.no_secondary:
        # BASELINE.json configs[1]: nano-GenCast, one member per GPU (launch / latency bound: 139 kernels of ~8 us)
        nano = measure(args, "nano", 1, max(args.steps, 3), 3, dev, rank, world, detailed=False)
        also = {"nano_2p5deg_one_member_per_gpu": {
            "workload": workload_name("nano"), "value": nano["value"], "unit": UNIT, "ms_per_step": nano["ms_per_step"],
            "denoiser_fwd_ms": nano["denoiser_fwd_ms"], "e2e": nano["e2e"], "members": world}}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline beside it (bounded sample; rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        case = main["case"]
        reps = 2 if args.config in ("nano", "tiny") else 1
        times = cpu_solver_iteration_seconds(case, repeats=reps, warmup=1 if args.config in ("nano", "tiny") else 0)
        v = 1.0 / (float(np.mean(times)) * 20)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
               "sample": f"{reps} of the 20 solver iterations (2 denoiser evaluations each) of one member, torch-fp32 oracle "
                         f"of the reference algorithm, all host threads; member-step time = 20 x mean iteration"}

    line = {"metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": workload_name(args.config), "members": world * MB, "members_per_gpu": MB,
                       "denoiser_evaluations_per_step": main["evals"],
                       "weights": "random N(0, 1/fan_in) (reference init makes the transformer an identity)",
                       "l2": "flushed between timed steps (256 MiB write, not timed)",
                       "execution": "one CUDA graph per 12 h step",
                       "collective": "nccl all_reduce of ensemble sum / sum-of-squares per step" if world > 1 else "none"},
            "denoiser_fwd_ms": main["denoiser_fwd_ms"], "denoiser_fwd_ms_per_member": main["denoiser_fwd_ms"] / MB,
            "denoiser_fwd_ms_eager_launch": main["denoiser_fwd_ms_eager_launch"],
            "e2e": main["e2e"],
            **({"stats_check": main["stats_check"], "crps": main["crps"]} if "stats_check" in main else {}),
            "gpu_launches": args.steps * main["launches_per_step"],
            "clocks": main["clocks"], "roofline": main["roofline"], "kernels": main["kernels"],
            "kernels_note": "per-launch device durations from CUDA-event pairs around each launch of 3 eagerly "
                            "replayed denoiser evaluations (queued behind a device delay); shares are of their sum"}
    if "rollout" in main:
        line["rollout"] = main["rollout"]
    if also is not None:
        line["also"] = also
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def roofline_clean(r):
    return {k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in r.items()}


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The single JSON line goes to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    # Libraries (NCCL's version banner, torchrun notices) write to fd 1: keep the real stdout for the
    # JSON line only and send everything else to stderr.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gpu", choices=["gpu", "reference"])
    ap.add_argument("--config", default="1deg", choices=["tiny", "nano", "1deg", "0p25deg"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--members-per-gpu", type=int, default=None,
                    help="ensemble members evaluated together on each GPU (default: 4 for 1deg = BASELINE configs[3]'s "
                         "32 members / 8 GPUs, 1 otherwise)")
    ap.add_argument("--no-secondary", action="store_true", help="skip the nano (configs[1]) measurement in the 1deg run")
    ap.add_argument("--rollout-steps", type=int, default=0,
                    help="also time an autoregressive rollout of this many 12 h steps (configs[1] / [3] name 30)")
    args = ap.parse_args()
    if args.members_per_gpu is None:
        args.members_per_gpu = DEFAULT_MEMBERS_PER_GPU.get(args.config, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_gpu(args)


if __name__ == "__main__":
    main()
