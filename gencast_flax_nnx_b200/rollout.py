"""`rollout.chunked_prediction` of the reference (common/rollout.py:205-401): the host
autoregressive driver around a PredictorFn(rng, inputs, targets_template, forcings)."""
from __future__ import annotations

from typing import Callable, Iterator

import numpy as np

from . import rngs as _rngs
from .xarray_lite import DataArray, Dataset, concat_time, merge

PredictorFn = Callable[..., Dataset]


def _with_time(ds: Dataset, time) -> Dataset:
    coords = dict(ds.coords)
    coords["time"] = np.asarray(time)
    return Dataset(ds.data_vars, coords)


def _get_next_inputs(prev_inputs: Dataset, next_frame: Dataset) -> Dataset:
    """Reference: common/rollout.py:379-401."""
    missing = set(prev_inputs.keys()) - set(next_frame.keys())
    for k in missing:
        if "time" in prev_inputs[k].dims:
            raise ValueError("Found an input with a time index that is not predicted or forced.")
    keys = [k for k in prev_inputs.keys() if k in next_frame]
    num_inputs = prev_inputs.sizes["time"]
    out = {}
    for k, v in prev_inputs.items():
        if k in keys and "time" in v.dims:
            nf = next_frame[k].transpose(*v.dims)
            ax = v.dims.index("time")
            cat = np.concatenate([v.data, nf.data], axis=ax)
            out[k] = DataArray(np.take(cat, range(cat.shape[ax] - num_inputs, cat.shape[ax]), axis=ax), v.dims)
        else:
            out[k] = v
    return Dataset(out, prev_inputs.coords)


def chunked_prediction_generator(predictor_fn: PredictorFn, rng, inputs: Dataset, targets_template: Dataset,
                                 forcings: Dataset, num_steps_per_chunk: int = 1, verbose: bool = False,
                                 pmap_devices=None) -> Iterator[Dataset]:
    """Reference: common/rollout.py:245-376 (same validation, same time re-labelling)."""
    if pmap_devices is not None:
        raise NotImplementedError("members are sharded one process per GPU (see parallel.py), not with pmap")
    num_target_steps = targets_template.sizes["time"]
    num_chunks, remainder = divmod(num_target_steps, num_steps_per_chunk)
    if remainder != 0:
        raise ValueError(f"The number of steps per chunk {num_steps_per_chunk} must "
                         f"evenly divide the number of target steps {num_target_steps} ")
    times = np.asarray(targets_template.coords["time"])
    if len(np.unique(np.diff(times))) > 1:
        raise ValueError("The targets time coordinates must be evenly spaced")
    targets_chunk_time = times[:num_steps_per_chunk]
    current_inputs = inputs
    for chunk_index in range(num_chunks):
        if verbose:
            print(f"Chunk {chunk_index}/{num_chunks}", flush=True)
        sl = slice(num_steps_per_chunk * chunk_index, num_steps_per_chunk * (chunk_index + 1))
        actual_time = times[sl]
        cur_t = _with_time(targets_template.isel(time=sl), targets_chunk_time)
        cur_f = _with_time(forcings.isel(time=sl), targets_chunk_time)
        rng, this_rng = _rngs.split(rng)
        predictions = predictor_fn(rng=this_rng, inputs=current_inputs, targets_template=cur_t, forcings=cur_f)
        next_frame = merge([predictions, cur_f])
        current_inputs = _with_time(_get_next_inputs(current_inputs, next_frame), current_inputs.coords["time"])
        yield _with_time(predictions, actual_time)


def chunked_prediction(predictor_fn: PredictorFn, rng, inputs: Dataset, targets_template: Dataset,
                       forcings: Dataset, num_steps_per_chunk: int = 1, verbose: bool = False) -> Dataset:
    """Reference: common/rollout.py:205-242."""
    chunks = list(chunked_prediction_generator(predictor_fn, rng, inputs, targets_template, forcings,
                                               num_steps_per_chunk, verbose))
    return concat_time(chunks)


def chunked_prediction_generator_multiple_runs(predictor_fn: PredictorFn, rngs, inputs: Dataset, targets_template: Dataset,
                                               forcings, num_samples: int, pmap_devices=None,
                                               **chunked_prediction_kwargs) -> Iterator[Dataset]:
    """Ensemble fan-out of the reference (common/rollout.py:78-202): one trajectory per sample, every yielded chunk
    labelled with its `sample` coordinate; `inputs` / `forcings` may carry a leading 'sample' dim (one initial
    condition per member) or be shared.  `rngs[i]` seeds sample i.

    Where the reference fans samples out with `pmap` over `pmap_devices` (:109-175), members here are sharded one
    process per GPU: with `pmap_devices` given (any sequence whose length is the number of GPUs / ranks), this process
    runs the samples `parallel.member_assignment(num_samples, len(pmap_devices), rank)` assigns to it -- the same
    divisibility rule as the reference (:110-112) -- and yields only those; without it every sample runs here, in order.
    """
    if num_samples is None:
        num_samples = len(rngs)
    if pmap_devices is not None:
        from . import parallel
        world = len(pmap_devices)
        if num_samples % world != 0:
            raise AssertionError("num_samples must be a multiple of len(pmap_devices)")
        rank, _ = parallel._world(None)
        samples = parallel.member_assignment(num_samples, world, rank % world)
    else:
        samples = range(num_samples)
    for i in samples:
        sample_inputs = inputs.isel(sample=i) if "sample" in inputs.sizes else inputs
        sample_forcings = forcings
        if sample_forcings is not None and "sample" in sample_forcings.sizes:
            sample_forcings = sample_forcings.isel(sample=i)
        if "sample" in sample_inputs.coords:
            sample_inputs = Dataset(sample_inputs.data_vars, {k: v for k, v in sample_inputs.coords.items() if k != "sample"})
        for chunk in chunked_prediction_generator(predictor_fn=predictor_fn, rng=rngs[i], inputs=sample_inputs,
                                                  targets_template=targets_template, forcings=sample_forcings,
                                                  **chunked_prediction_kwargs):
            coords = dict(chunk.coords)
            coords["sample"] = np.asarray(i)
            yield Dataset(chunk.data_vars, coords)


# ----------------------------------------------------------------------------------------------
# The same rollout with the autoregressive window resident on the GPU
# ----------------------------------------------------------------------------------------------

_PRESERVED = ("batch", "lat", "lon")


def _channel_index(var: DataArray):
    """Integer array over the variable's non-(batch, lat, lon) dims (in its own dim order) holding the channel
    number each element is stacked to (stacking.variable_to_nodes: row-major over those dims)."""
    extra = [d for d in var.dims if d not in _PRESERVED]
    shape = [var.sizes[d] for d in extra]
    return extra, np.arange(int(np.prod(shape, dtype=np.int64)) if shape else 1, dtype=np.int64).reshape(shape or ())


def window_update_table(inputs: Dataset, predictions: Dataset, forcings: Dataset) -> np.ndarray:
    """Column table of the on-device window update (gc_select_columns): entry j says where channel j of the
    next step's stacked inputs comes from -- (0, c) the current stacked inputs, (1, c) the stacked prediction,
    (2, c) the stacked forcings of the step -- encoded as source << 24 | c.

    Semantics of the reference's `_get_next_inputs` (common/rollout.py:379-401) with a one-step chunk: for every
    input variable with a time axis, frames shift by one and the last frame is the variable's value in
    merge([predictions, forcings]); variables without a time axis are kept; an input with a time axis that is
    neither predicted nor forced is an error."""
    def offsets(ds):
        out, off = {}, 0
        for n in sorted(ds.keys()):
            _, idx = _channel_index(ds[n])
            out[n] = off
            off += idx.size
        return out, off
    in_off, n_in = offsets(inputs)
    pr_off, _ = offsets(predictions)
    fr_off, _ = offsets(forcings)
    table = np.zeros(n_in, np.int64)
    for name in sorted(inputs.keys()):
        v = inputs[name]
        extra, idx = _channel_index(v)
        base = in_off[name]
        if "time" not in extra:
            table[base + idx.reshape(-1)] = (0 << 24) | (base + idx.reshape(-1))
            continue
        if name in predictions:
            src_id, src_var, src_base = 1, predictions[name], pr_off[name]
        elif name in forcings:
            src_id, src_var, src_base = 2, forcings[name], fr_off[name]
        else:
            raise ValueError("Found an input with a time index that is not predicted or forced.")
        s_extra, s_idx = _channel_index(src_var)
        if "time" not in s_extra or src_var.sizes["time"] != 1:
            raise ValueError("the on-device rollout advances one target step at a time")
        # bring the source's extra dims into the input variable's order, time first dropped
        s_idx = np.transpose(s_idx, [s_extra.index(d) for d in extra])
        ax = extra.index("time")
        T = v.sizes["time"]
        older = np.take(idx, range(1, T), axis=ax)                  # frames 1 .. T-1 move to 0 .. T-2
        table[base + np.take(idx, range(0, T - 1), axis=ax).reshape(-1)] = (0 << 24) | (base + older.reshape(-1))
        table[base + np.take(idx, [T - 1], axis=ax).reshape(-1)] = (src_id << 24) | (src_base + s_idx.reshape(-1))
    return table.astype(np.int32)


def _wrapper_transforms(model, inputs: Dataset, targets_template: Dataset, forcings: Dataset):
    """Unwraps `normalization.InputsAndResiduals` / `nan_cleaning.NaNCleaner` around a GenCast model (in either order,
    as training/train_helpers.py:160-214 composes them) into per-stacked-channel vectors for gc_normalize_cast /
    gc_unnormalize_residual.  Returns (gencast model, dict of numpy vectors or None)."""
    from .nan_cleaning import NaNCleaner
    from .normalization import InputsAndResiduals
    chain = []
    while hasattr(model, "predictor"):
        chain.append(model)
        model = model.predictor
    if not chain:
        return model, None
    t = {}
    normalised = False
    for w in chain:                                        # outermost first: the order the inputs pass through
        if isinstance(w, InputsAndResiduals):
            if normalised:
                raise ValueError("two InputsAndResiduals wrappers")
            t.update(w.channel_transforms(inputs, targets_template, forcings))
            # the residual is added to the inputs AS THIS WRAPPER SEES THEM: cleaned if a NaNCleaner sits outside
            if "in_fill_pre" in t:
                fill_in = t["in_fill_pre"]
                t["res_fill"] = np.where(t["res_col"] >= 0, fill_in[np.maximum(t["res_col"], 0)], np.nan).astype(np.float32)
            normalised = True
        elif isinstance(w, NaNCleaner):
            side = "post" if normalised else "pre"
            t[f"in_fill_{side}"] = w.channel_fill(inputs)
            t[f"frc_fill_{side}"] = w.channel_fill(forcings)
            if w._reintroduce_nans and w._var_to_clean in targets_template and w._var_to_clean in inputs:
                in_off, off = {}, 0
                for n in sorted(inputs.keys()):
                    in_off[n] = off
                    off += _channel_index(inputs[n])[1].size
                iv = inputs[w._var_to_clean]
                i_extra, i_idx = _channel_index(iv)
                T = iv.sizes["time"]
                cols = []
                for n in sorted(targets_template.keys()):
                    extra, idx = _channel_index(targets_template[n])
                    if n != w._var_to_clean:
                        cols.append(np.full((idx.size, T), -1, np.int64))
                        continue
                    frames = [np.transpose(np.take(i_idx, [k], axis=i_extra.index("time")), [i_extra.index(d) for d in extra]).reshape(-1)
                              for k in range(T)]
                    cols.append(in_off[n] + np.stack(frames, axis=1))
                t["nan_cols"] = np.concatenate(cols).astype(np.int32)
        else:
            raise TypeError(f"device rollout: unsupported predictor wrapper {type(w).__name__}")
    return model, t


def device_chunked_prediction_generator(model, inputs: Dataset, targets_template: Dataset, forcings: Dataset,
                                        verbose: bool = False) -> Iterator[Dataset]:
    """`chunked_prediction_generator` (common/rollout.py:245-376) for a `gencast.GenCast` model with one target
    step per chunk, keeping the autoregressive input window on the GPU: the inputs go host -> device once,
    each step uploads only its forcings, the next window is assembled on the device from the previous window,
    the prediction and the forcings (`window_update_table`, gc_select_columns), and only the prediction comes
    back to the host.  Step for step it yields what
    `chunked_prediction_generator(lambda rng, inputs, targets_template, forcings: model.full_sampling(inputs,
    targets_template, forcings), ...)` yields with the same `model.rngs` state.

    `model` may be wrapped in `normalization.InputsAndResiduals` and / or `nan_cleaning.NaNCleaner`: the window then holds
    physical values, gc_normalize_cast normalises / cleans them into the network's operand every step and
    gc_unnormalize_residual turns the sample back into physical units with the residual connection (and the NaNs put
    back where asked) before it enters the window -- the reference's wrappers without leaving the device."""
    import torch
    from . import ops
    outer = model
    model, tr_np = _wrapper_transforms(outer, inputs, targets_template.isel(time=slice(0, 1)), forcings.isel(time=slice(0, 1)))
    tr = None
    if model._sampler is None:
        raise ValueError("Sampler config must be specified to run inference.")
    times = np.asarray(targets_template.coords["time"])
    if len(np.unique(np.diff(times))) > 1:
        raise ValueError("The targets time coordinates must be evenly spaced")
    den, sampler = model.denoiser, model._sampler
    chunk_time = times[:1]
    window = None
    table = None
    for step in range(targets_template.sizes["time"]):
        if verbose:
            print(f"Chunk {step}/{targets_template.sizes['time']}", flush=True)
        sl = slice(step, step + 1)
        cur_t = _with_time(targets_template.isel(time=sl), chunk_time)
        cur_f = _with_time(forcings.isel(time=sl), chunk_time)
        engine = den._maybe_init(inputs, cur_t, cur_f)
        sizes = dict(cur_t.sizes)
        sizes.setdefault("batch", 1)
        with torch.cuda.device(engine.device):
            if window is None:
                window = den.member_major(den.stacker.to_nodes("inputs", inputs, sizes)).clone()      # [B*G, C_in] fp32
                nxt = torch.empty_like(window)
                table = torch.from_numpy(window_update_table(inputs, cur_t, cur_f)).to(engine.device)
                if tr_np is not None:
                    tr = {k: torch.from_numpy(np.ascontiguousarray(v)).to(engine.device) for k, v in tr_np.items()}
                    phys = torch.empty(window.shape[0], engine.n_out, dtype=torch.float32, device=engine.device)
            frc = den.member_major(den.stacker.to_nodes("forcings", cur_f, sizes))
            engine.set_constant_features(window, frc, transform=tr)
            res = sampler.sample_on_device(cur_t, model.rngs.noise())
            if tr is not None:
                res = ops.unnormalize_residual(res, engine.n_out, phys, tr.get("out_scale"), tr.get("out_loc"), window,
                                               tr.get("res_col"), tr.get("res_fill"), tr.get("nan_cols"))
            out = res.reshape(sizes["batch"], engine.G, engine.n_out).permute(1, 0, 2)
            pred = den.stacker.from_nodes(out, cur_t)
            ops.select_columns([window, res, frc], table, nxt)
            window, nxt = nxt, window
        yield _with_time(pred, times[sl])


def device_chunked_prediction(model, inputs: Dataset, targets_template: Dataset, forcings: Dataset,
                              verbose: bool = False) -> Dataset:
    return concat_time(list(device_chunked_prediction_generator(model, inputs, targets_template, forcings, verbose)))
