"""Dataset <-> [grid node, batch, channel] layout.

Restates, for the xarray_lite containers, the reference's
common/model_utils.py:594-725 (variable_to_stacked, dataset_to_stacked,
stacked_to_dataset) and gencast/denoiser.py:770-830: variables are taken in
sorted-name order; every dim other than (batch, lat, lon) is folded into
channels in the variable's own dim order (time-major for (time, level));
missing batch/lat/lon dims are broadcast; node index = lat_idx * n_lon + lon_idx.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Tuple

import numpy as np

from .xarray_lite import DataArray, Dataset

PRESERVED = ("batch", "lat", "lon")


def variable_channels(var: DataArray) -> int:
    n = 1
    for d, s in var.sizes.items():
        if d not in PRESERVED:
            n *= s
    return n


def variable_to_nodes(var: DataArray, sizes: Mapping[str, int]) -> np.ndarray:
    """One variable -> [lat*lon, batch, channels] (reference: model_utils.py:594-623)."""
    extra = [d for d in var.dims if d not in PRESERVED]
    have = [d for d in PRESERVED if d in var.dims]
    arr = var.transpose(*(have + extra)).data
    arr = arr.reshape(arr.shape[:len(have)] + (-1,))
    # insert missing preserved dims as size-1 axes, then broadcast
    shape, k = [], 0
    for d in PRESERVED:
        if d in have:
            shape.append(arr.shape[k]); k += 1
        else:
            shape.append(1)
    arr = arr.reshape(tuple(shape) + (arr.shape[-1],))
    full = (sizes["batch"], sizes["lat"], sizes["lon"], arr.shape[-1])
    arr = np.broadcast_to(arr, full)
    # (batch, lat, lon, c) -> (lat, lon, batch, c) -> (node, batch, c)   (denoiser.py:801-806)
    return np.ascontiguousarray(np.transpose(arr, (1, 2, 0, 3))).reshape(full[1] * full[2], full[0], full[3])


def dataset_to_nodes(ds: Dataset, sizes: Mapping[str, int]) -> Tuple[np.ndarray, List[Tuple[str, int]]]:
    """Sorted-name concat of all variables -> ([node, batch, C], [(name, channels), ...])."""
    names = sorted(ds.keys())
    if not names:
        return np.zeros((sizes["lat"] * sizes["lon"], sizes["batch"], 0), np.float32), []
    blocks = [variable_to_nodes(ds[n], sizes) for n in names]
    return np.concatenate(blocks, axis=-1), [(n, b.shape[-1]) for n, b in zip(names, blocks)]


def channel_layout(template: Dataset) -> List[Tuple[str, int]]:
    """[(name, channels)] in stacking order for a template dataset."""
    return [(n, variable_channels(template[n])) for n in sorted(template.keys())]


def nodes_to_dataset(nodes: np.ndarray, template: Dataset) -> Dataset:
    """Inverse of dataset_to_nodes for a template whose variables all carry batch/lat/lon.

    Reference: gencast/denoiser.py:809-830 and common/model_utils.py:662-725.
    nodes: [lat*lon, batch, C].
    """
    sizes = template.sizes
    n_lat, n_lon = sizes["lat"], sizes["lon"]
    out: Dict[str, DataArray] = {}
    i = 0
    for name in sorted(template.keys()):
        tv = template[name]
        if not all(d in tv.dims for d in PRESERVED):
            raise ValueError(f"stacked_to_dataset requires all variables to have {PRESERVED} dimensions, "
                             f"but found only {tv.dims}.")
        extra = [d for d in tv.dims if d not in PRESERVED]
        extra_shape = [tv.sizes[d] for d in extra]
        c = int(np.prod(extra_shape, dtype=np.int64)) if extra else 1
        blk = nodes[:, :, i:i + c]
        i += c
        blk = blk.reshape([n_lat, n_lon, blk.shape[1]] + extra_shape)          # lat, lon, batch, extra...
        cur = ["lat", "lon", "batch"] + extra
        out[name] = DataArray(blk, cur).transpose(*tv.dims)
    if i != nodes.shape[-1]:
        raise ValueError(f"Expected {i} channels but found {nodes.shape[-1]}")
    return Dataset(out, template.coords)
