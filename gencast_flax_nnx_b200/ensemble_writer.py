"""Packed on-disk layout for ensemble rollouts: one flat file of [time, member, grid node, channel] values plus a JSON
header, written step by step straight from the sampler's state layout.

The reference persists a rollout with xarray's netCDF writer, one compressed variable at a time after gathering the
whole forecast on the host (`save_rollout_to_netcdf`, training/evaluation.py:194-266).  A 15-day, 32-member, 1 deg
forecast is 30 x 32 x 65 160 x 82 floats = 20.5 GB; here each 12 h step of the members resident on a GPU is one
contiguous [members, G, 82] block in exactly the layout the sampler produces ([member * G + node, channel]), so writing
is a device -> pinned-host copy and one sequential file write, and ranks write disjoint member ranges of the same file.
`PackedRolloutReader` maps the file and unstacks blocks back into per-variable Datasets (stacking.nodes_to_dataset,
i.e. the reference's stacked_to_dataset, common/model_utils.py:662-725)."""
from __future__ import annotations

import json
import os
from typing import Optional, Sequence

import numpy as np

from . import stacking
from .xarray_lite import Dataset

MAGIC = "gencast-b200-packed-rollout-v1"


class PackedRolloutWriter:
    def __init__(self, path: str, template: Dataset, times: Sequence, num_members: int, dtype: str = "float32",
                 attrs: Optional[dict] = None, create: bool = True):
        """template: one-step targets template (variables / dims / lat / lon define the channel layout)."""
        sizes = template.sizes
        self.G = sizes["lat"] * sizes["lon"]
        self.layout = stacking.channel_layout(template)
        self.C = sum(c for _, c in self.layout)
        self.T, self.M = len(times), int(num_members)
        self.dtype = np.dtype(dtype)
        self.path = path
        header = dict(magic=MAGIC, dtype=self.dtype.name, shape=[self.T, self.M, self.G, self.C],
                      times=[float(t) for t in np.asarray(times).reshape(-1)], lat=np.asarray(template.coords["lat"], float).tolist(),
                      lon=np.asarray(template.coords["lon"], float).tolist(),
                      variables=[[n, c, list(template[n].dims), [int(s) for s in template[n].shape]] for n, c in self.layout],
                      coords={k: np.asarray(v, float).tolist() for k, v in template.coords.items() if k not in ("lat", "lon", "time", "batch")},
                      attrs=attrs or {})
        if create:
            with open(path + ".json", "w") as f:
                json.dump(header, f)
            with open(path, "wb") as f:
                f.truncate(self.T * self.M * self.G * self.C * self.dtype.itemsize)
        self._mm = np.memmap(path, dtype=self.dtype, mode="r+", shape=(self.T, self.M, self.G, self.C))
        self._pin = None

    def write_step(self, step: int, block, first_member: int = 0) -> None:
        """block: [members * G, C] (the sampler's member-major state layout) as a host array or a CUDA tensor."""
        if hasattr(block, "is_cuda") and block.is_cuda:
            import torch
            if self._pin is None or self._pin.shape != block.shape:
                self._pin = torch.empty(block.shape, dtype=torch.float32, pin_memory=True)
            self._pin.copy_(block, non_blocking=True)
            torch.cuda.current_stream(block.device).synchronize()
            block = self._pin.numpy()
        block = np.asarray(block)
        m = block.shape[0] // self.G
        if block.shape != (m * self.G, self.C) or first_member + m > self.M:
            raise ValueError(f"block of shape {block.shape} does not fit members {first_member}.. of {self.M}")
        self._mm[step, first_member:first_member + m] = block.reshape(m, self.G, self.C).astype(self.dtype, copy=False)

    def close(self) -> None:
        self._mm.flush()
        del self._mm


class PackedRolloutReader:
    def __init__(self, path: str):
        with open(path + ".json") as f:
            self.header = h = json.load(f)
        if h.get("magic") != MAGIC:
            raise ValueError("not a packed rollout file")
        self.shape = tuple(h["shape"])
        self.data = np.memmap(path, dtype=np.dtype(h["dtype"]), mode="r", shape=self.shape)

    def template(self, batch: int = 1) -> Dataset:
        from .xarray_lite import DataArray
        h = self.header
        coords = dict(lat=np.asarray(h["lat"], np.float32), lon=np.asarray(h["lon"], np.float32),
                      **{k: np.asarray(v) for k, v in h["coords"].items()})
        out = {}
        for name, _, dims, shape in h["variables"]:
            shape = [batch if d == "batch" else s for d, s in zip(dims, shape)]
            out[name] = DataArray(np.zeros(shape, np.float32), dims)
        return Dataset(out, coords)

    def step(self, step: int, members: Optional[slice] = None) -> Dataset:
        """One forecast step of the selected members as a Dataset with batch = members."""
        blk = np.asarray(self.data[step, members if members is not None else slice(None)], np.float32)    # [m, G, C]
        ds = stacking.nodes_to_dataset(np.ascontiguousarray(np.transpose(blk, (1, 0, 2))), self.template(blk.shape[0]))
        coords = dict(ds.coords)
        coords["time"] = np.asarray([self.header["times"][step]])
        return Dataset(ds.data_vars, coords)
