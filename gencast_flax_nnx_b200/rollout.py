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
