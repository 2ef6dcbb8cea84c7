"""Seeded synthetic ERA5-shaped inputs for tests and benchmarks.

Variable sets and dims follow gencast.TASK (reference: gencast/gencast.py:57-71):
inputs carry 2 time steps of 4 surface + 6x13 atmospheric variables, the four
progress features and two static fields; forcings carry the progress features
for the target step; targets are one step of the 82 predicted channels.
Values are standard normal (data is assumed normalised, reference:
common/normalization.py:155).
"""
from __future__ import annotations

import numpy as np

from .configs import (ALL_ATMOSPHERIC_VARS, GENERATED_FORCING_VARS, STATIC_VARS, TASK, TaskConfig)
from .xarray_lite import DataArray, Dataset


def make_example(grid_lat, grid_lon, batch: int = 1, seed: int = 0, task: TaskConfig = TASK,
                 num_target_steps: int = 1):
    """Returns (inputs, targets_template, forcings) Datasets."""
    rng = np.random.default_rng(seed)
    n_lat, n_lon = len(grid_lat), len(grid_lon)
    n_lev = len(task.pressure_levels)
    coords = dict(lat=np.asarray(grid_lat), lon=np.asarray(grid_lon), level=np.asarray(task.pressure_levels),
                  batch=np.arange(batch))
    f32 = np.float32

    def field(n_time, levels):
        shape = (batch, n_time, n_lev, n_lat, n_lon) if levels else (batch, n_time, n_lat, n_lon)
        dims = ("batch", "time", "level", "lat", "lon") if levels else ("batch", "time", "lat", "lon")
        return DataArray(rng.standard_normal(shape).astype(f32), dims)

    inputs, targets = {}, {}
    for name in task.input_variables:
        if name in STATIC_VARS:
            inputs[name] = DataArray(rng.standard_normal((n_lat, n_lon)).astype(f32), ("lat", "lon"))
        elif name in GENERATED_FORCING_VARS:
            phase = rng.uniform(0, 2 * np.pi, size=(batch, 2)).astype(f32)
            fn = np.sin if name.endswith("sin") else np.cos
            if name.startswith("day"):
                lon_phase = np.deg2rad(np.asarray(grid_lon, dtype=f32))
                inputs[name] = DataArray(fn(phase[:, :, None] + lon_phase[None, None, :]).astype(f32),
                                         ("batch", "time", "lon"))
            else:
                inputs[name] = DataArray(fn(phase).astype(f32), ("batch", "time"))
        else:
            inputs[name] = field(2, name in ALL_ATMOSPHERIC_VARS)
    for name in task.target_variables:
        levels = name in ALL_ATMOSPHERIC_VARS
        shape = ((batch, num_target_steps, n_lev, n_lat, n_lon) if levels
                 else (batch, num_target_steps, n_lat, n_lon))
        dims = ("batch", "time", "level", "lat", "lon") if levels else ("batch", "time", "lat", "lon")
        targets[name] = DataArray(np.zeros(shape, f32), dims)
    forcings = {}
    for name in task.forcing_variables:
        phase = rng.uniform(0, 2 * np.pi, size=(batch, num_target_steps)).astype(f32)
        fn = np.sin if name.endswith("sin") else np.cos
        if name.startswith("day"):
            lon_phase = np.deg2rad(np.asarray(grid_lon, dtype=f32))
            forcings[name] = DataArray(fn(phase[:, :, None] + lon_phase[None, None, :]).astype(f32),
                                       ("batch", "time", "lon"))
        else:
            forcings[name] = DataArray(fn(phase).astype(f32), ("batch", "time"))
    c_in = dict(coords, time=np.arange(-1, 1) * 12)
    c_tg = dict(coords, time=(np.arange(num_target_steps) + 1) * 12)
    return Dataset(inputs, c_in), Dataset(targets, c_tg), Dataset(forcings, c_tg)
