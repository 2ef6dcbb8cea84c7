#!/usr/bin/env python
"""Generates tests/golden/refshim_sampler.npz by executing the REFERENCE'S OWN code for the parts of the hot path
that tools/make_refshim_golden.py does not reach:

  * `FourierFeaturesMLP.__call__` (common/mlp.py:207-265) with `fourier_features` (common/model_utils.py:728-757),
    constructed as gencast/denoiser.py:169-170 constructs the noise-level encoder;
  * `Sampler.__call__` with its `body_fn`, `denoise_arr` and `_preconditioned_denoiser`
    (gencast/dpm_solver_plus_plus_2s.py:47-205), `noise_schedule` / `stochastic_churn_rate_schedule`
    (gencast/samplers_utils.py:395-431), run unmodified around a small deterministic stand-in network;
  * the Dataset <-> [node, batch, channel] stacking (common/model_utils.py:145-167, :594-725) as
    gencast/denoiser.py:184, :770-830 calls it.

Third-party packages are replaced by numpy stand-ins (tools/refshim/: jax, flax.nnx, chex, xarray); two of the
reference's own modules cannot be imported and are replaced by the few functions used from them:
`common/xarray_jax.py` (JAX pytree registration of xarray types: `unwrap`, `unwrap_data`, `DataArray`) and the
dinosaur-based noise generator `samplers_utils.spherical_white_noise_like` (dinosaur is absent; the initial noise is
an input of this fixture instead).  Only runnable where /root/reference exists; the fixture travels with the
repository and tests/test_oracle_golden.py checks the oracle, the host stacking code and (on the GPU) gc_cond_tables
against it.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("GENCAST_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

SIGMAS = (80.0, 7.5, 1.0, 0.03, 1e-6)
NUM_LEVELS = 5
GRID_RES = 30.0
BATCH = 2


def _install_shims():
    shim = os.path.join(ROOT, "tools", "refshim")
    for p in (REFERENCE, shim):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE)
    sys.path.insert(0, shim)
    import xarray as xr                                     # the stand-in
    if "common.xarray_jax" not in sys.modules:
        import common                                       # the reference's package
        xj = types.ModuleType("common.xarray_jax")
        xj.unwrap = lambda v, require_jax=False: v.data if isinstance(v, xr.Variable) else v
        xj.unwrap_data = xj.unwrap
        xj.jax_data = xj.unwrap_data
        xj.DataArray = xr.DataArray
        xj.Variable = xr.Variable
        xj.Dataset = xr.Dataset
        sys.modules["common.xarray_jax"] = xj
        common.xarray_jax = xj
    for name in ("dinosaur", "dinosaur.spherical_harmonic"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sh = sys.modules["dinosaur.spherical_harmonic"]
    sh.Grid = sh.RealSphericalHarmonics = type("Absent", (), {})      # names used in annotations only
    sys.modules["dinosaur"].spherical_harmonic = sh
    return xr


def toy_weights(c_in: int, n_out: int):
    rng = np.random.default_rng(5)
    return rng.standard_normal((c_in, n_out)) / np.sqrt(c_in)


def toy_network(feats, sigma, w):
    """Stand-in for the GenCast network: [G, B, C] features, [B] noise levels -> [G, B, n_out]."""
    return np.tanh(feats @ w) * np.cos(np.log(sigma))[None, :, None] + 0.25 * np.sin(feats[..., :1])


def build_case():
    """Seeded Datasets (xarray_lite) on a 30 deg grid; target variables in sorted order (the
    reference's sampler mixes variables when the template is not name-sorted)."""
    from gencast_flax_nnx_b200 import graph, synthetic
    lat, lon = graph.regular_grid(GRID_RES)
    inputs, targets, forcings = synthetic.make_example(lat, lon, batch=BATCH, seed=3)
    rng = np.random.default_rng(17)
    noise = {k: rng.standard_normal(v.shape) for k, v in sorted(targets.items())}
    return inputs, targets, forcings, noise


def to_shim(xr, ds, order=None, dtype=np.float64):
    names = list(ds.keys()) if order is None else order
    coords = {k: np.asarray(v) for k, v in ds.coords.items()}
    out = {}
    for n in names:
        v = ds[n]
        out[n] = xr.DataArray(np.asarray(v.data, dtype), coords={d: coords[d] for d in v.dims if d in coords}, dims=v.dims, name=n)
    return xr.Dataset(out, coords={k: v for k, v in coords.items() if any(k in out[n].dims for n in out)})


def run_reference():
    xr = _install_shims()
    import flax.nnx as nnx
    from common import mlp as ref_mlp
    from common import model_utils as mu
    from common import xarray_jax
    try:
        from graphcast import casting  # noqa: F401  (the reference's own infer_floating_dtype)
    except Exception as e:                                  # pragma: no cover
        raise RuntimeError(f"graphcast.casting does not import under the stand-ins: {e}")
    from gencast import samplers_utils as su
    from gencast import dpm_solver_plus_plus_2s as ref_sampler
    from gencast_flax_nnx_b200 import params as our_params, configs

    out = {}
    # ---------------- FourierFeaturesMLP, as gencast/denoiser.py:166-170 builds it (NoiseEncoderConfig defaults :57-61)
    enc = ref_mlp.FourierFeaturesMLP(apply_log_first=True, base_period=16.0, num_frequencies=32, output_sizes=(32, 16),
                                     rngs=nnx.Rngs(0))
    _, arch = configs.named_config("tiny")
    p = our_params.init_perturbed(our_params.param_shapes(arch, 20, 7), seed=1)
    pre = "denoiser/noise_level_encoder"
    seen = 0
    for path, prm in enc.named_params(pre):
        if "/linears/" in path:               # the module also keeps its layers in a list: same Param objects
            continue
        assert p[path].shape == prm.value.shape, path
        prm.value = np.asarray(p[path], np.float64)
        seen += 1
    assert seen == 4 == len([k for k in p if k.startswith(pre + "/")])
    out["encoder/sigmas"] = np.asarray(SIGMAS, np.float64)
    out["encoder/cond"] = np.asarray(enc(np.asarray(SIGMAS, np.float64)))
    for k in ("linear_0/kernel", "linear_0/bias", "linear_1/kernel", "linear_1/bias"):
        out[f"encoder/{k}"] = np.asarray(p[f"{pre}/{k}"], np.float64)

    # ---------------- stacking + sampler around a toy network
    inputs_l, targets_l, forcings_l, noise = build_case()
    tnames = sorted(targets_l.keys())
    inputs, forcings = to_shim(xr, inputs_l), to_shim(xr, forcings_l)
    targets = to_shim(xr, targets_l, order=tnames)
    noise_ds = xr.Dataset({n: xr.DataArray(noise[n], coords=targets[n].coords, dims=targets[n].dims, name=n) for n in tnames},
                          coords=targets.coords)

    class ToyDenoiser:
        """Data path of Denoiser.__call__ / DenoiserArchitecture (gencast/denoiser.py:172-202, :303-341, :770-830)
        with the three GNNs replaced by toy_network."""
        w = None

        def __call__(self, inputs, noisy_targets, noise_levels, forcings=None, **kw):
            if forcings is None:
                forcings = xr.Dataset()
            forcings = forcings.assign(noisy_targets)                               # :184
            if noise_levels.dims != ("batch",):                                     # :188-189
                raise ValueError("noise_levels expected to be shape (batch,).")
            stacked_inputs = mu.dataset_to_stacked(inputs)                          # :794
            stacked_forcings = mu.dataset_to_stacked(forcings)                      # :795
            stacked = xr.concat([stacked_inputs, stacked_forcings], dim="channels")  # :796-797
            lead = mu.lat_lon_to_leading_axes(stacked)                              # :801-802
            feats = xarray_jax.unwrap(lead.data).reshape((-1,) + lead.data.shape[2:])   # :804-806
            if ToyDenoiser.w is None:
                ToyDenoiser.w = toy_weights(feats.shape[-1], sum(int(np.prod([s for d, s in noisy_targets[n].sizes.items()
                                                                              if d not in ("batch", "lat", "lon")]))
                                                                 for n in noisy_targets.keys()))
                out["sampler/first_call_features"] = np.array(feats)
            raw = toy_network(feats, np.asarray(noise_levels.data, np.float64), ToyDenoiser.w)
            n_lat, n_lon = len(inputs.coords["lat"].data), len(inputs.coords["lon"].data)
            grid = raw.reshape((n_lat, n_lon) + raw.shape[1:])                      # :818-821
            da = xarray_jax.DataArray(data=grid, dims=("lat", "lon", "batch", "channels"))
            return mu.stacked_to_dataset(mu.restore_leading_axes(da).variable, noisy_targets)   # :825-830

    su.spherical_white_noise_like = lambda template, rngs: noise_ds      # dinosaur is absent: noise is an input
    ref_sampler.utils.spherical_white_noise_like = su.spherical_white_noise_like
    sampler = ref_sampler.Sampler(ToyDenoiser(), max_noise_level=80.0, min_noise_level=0.03, num_noise_levels=NUM_LEVELS,
                                  rho=7.0, stochastic_churn_rate=0.0, churn_min_noise_level=0.75,
                                  churn_max_noise_level=float("inf"), noise_level_inflation_factor=1.05)
    try:
        sampler(inputs, targets, forcings)
        raise AssertionError("the reference raises without rngs")
    except ValueError:
        pass
    result = sampler(inputs, targets, forcings, rngs=nnx.Rngs(0))
    out["sampler/noise_levels"] = np.asarray(sampler._noise_levels, np.float64)
    out["sampler/churn_rates"] = np.asarray(sampler._per_step_churn_rates, np.float64)
    # the reference's own churn-rate schedule for a non-zero rate (gencast/samplers_utils.py:414-431)
    out["sampler/churn_schedule_rate2p5"] = np.asarray(su.stochastic_churn_rate_schedule(
        sampler._noise_levels, 2.5, 0.75, float("inf")), np.float64)
    out["sampler/toy_w"] = ToyDenoiser.w
    for n in tnames:
        # dims come back in to_array()'s broadcast order (level last); xarray semantics are by name
        assert sorted(result[n].dims) == sorted(targets[n].dims), (n, result[n].dims, targets[n].dims)
        out[f"sampler/result/{n}"] = np.asarray(result[n].transpose(*targets[n].dims).data)
        out[f"sampler/noise/{n}"] = noise[n]
    # one preconditioned call on its own (dpm...py:190-205), sigma differing per batch element
    sig = xr.DataArray(np.asarray([3.0, 0.2]), coords={"batch": targets.coords["batch"]}, dims=("batch",))
    noisy = xr.Dataset({n: xr.DataArray(noise[n] * 2.0, coords=targets[n].coords, dims=targets[n].dims, name=n) for n in tnames},
                       coords=targets.coords)
    d = sampler._preconditioned_denoiser(inputs=inputs, noisy_targets=noisy, noise_levels=sig, forcings=forcings)
    for n in tnames:
        out[f"precond/result/{n}"] = np.asarray(d[n].transpose(*targets[n].dims).data)
    # stacking on its own: Dataset -> [node, batch, channels] and back
    st = mu.lat_lon_to_leading_axes(mu.dataset_to_stacked(inputs))
    out["stacking/inputs_nodes"] = np.asarray(st.data).reshape((-1,) + st.data.shape[2:])
    out["stacking/input_names_sorted"] = np.asarray(sorted(inputs.keys()))
    return out


def main():
    out = run_reference()
    dst = os.path.join(ROOT, "tests", "golden", "refshim_sampler.npz")
    np.savez_compressed(dst, **out)
    print(f"wrote {dst}: {os.path.getsize(dst) / 1e6:.2f} MB, {len(out)} arrays")


if __name__ == "__main__":
    main()
