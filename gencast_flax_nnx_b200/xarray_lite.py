"""A very small labelled-array container standing in for xarray at the API edge.

The reference passes `xarray.Dataset`s through its public calls
(gencast/gencast.py:289-294, gencast/denoiser.py:172-202,
common/rollout.py:205-213).  xarray is not installable in the build image, so
the mirror API accepts this minimal Dataset (dict of name -> DataArray(dims,
data)); `from_xarray` / `to_xarray` convert when the real package is present.
Only what the hot path needs is implemented: dims/sizes bookkeeping, isel along
one dim, assign/merge, and time concatenation for the rollout window.
"""
from __future__ import annotations

from typing import Dict, Iterable, Mapping, Optional, Sequence, Tuple

import numpy as np


class DataArray:
    """Array with named dimensions.  `data` is a numpy array (host side)."""

    __slots__ = ("data", "dims")

    def __init__(self, data, dims: Sequence[str]):
        data = np.asarray(data)
        dims = tuple(dims)
        if data.ndim != len(dims):
            raise ValueError(f"data has {data.ndim} axes but dims={dims}")
        self.data = data
        self.dims = dims

    @property
    def sizes(self) -> Dict[str, int]:
        return dict(zip(self.dims, self.data.shape))

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    def isel(self, **indexers) -> "DataArray":
        data, dims = self.data, list(self.dims)
        for dim, idx in indexers.items():
            if dim not in dims:
                continue
            ax = dims.index(dim)
            if isinstance(idx, slice):
                data = data[(slice(None),) * ax + (idx,)]
            else:
                data = np.take(data, idx, axis=ax)
                dims.pop(ax)
        return DataArray(data, dims)

    def transpose(self, *dims) -> "DataArray":
        perm = [self.dims.index(d) for d in dims]
        return DataArray(np.transpose(self.data, perm), dims)

    def __mul__(self, other):
        return DataArray(self.data * (other.data if isinstance(other, DataArray) else other), self.dims)

    __rmul__ = __mul__

    def __repr__(self):
        return f"DataArray(dims={self.dims}, shape={self.data.shape}, dtype={self.data.dtype})"


class Dataset:
    """Ordered mapping name -> DataArray with shared coordinates."""

    def __init__(self, data_vars: Optional[Mapping[str, DataArray]] = None,
                 coords: Optional[Mapping[str, np.ndarray]] = None):
        self.data_vars: Dict[str, DataArray] = dict(data_vars or {})
        self.coords: Dict[str, np.ndarray] = {k: np.asarray(v) for k, v in (coords or {}).items()}

    # mapping protocol
    def __getitem__(self, key):
        if isinstance(key, (list, tuple)):
            return Dataset({k: self.data_vars[k] for k in key}, self.coords)
        return self.data_vars[key]

    def __contains__(self, key):
        return key in self.data_vars

    def __iter__(self):
        return iter(self.data_vars)

    def keys(self):
        return self.data_vars.keys()

    def items(self):
        return self.data_vars.items()

    def __len__(self):
        return len(self.data_vars)

    @property
    def sizes(self) -> Dict[str, int]:
        out: Dict[str, int] = {}
        for v in self.data_vars.values():
            for d, n in v.sizes.items():
                if out.setdefault(d, n) != n:
                    raise ValueError(f"inconsistent size for dim {d!r}")
        return out

    dims = sizes

    @property
    def lat(self):
        return self.coords["lat"]

    @property
    def lon(self):
        return self.coords["lon"]

    def assign(self, other=None, **kw) -> "Dataset":
        new = dict(self.data_vars)
        if other is not None:
            new.update(other.data_vars if isinstance(other, Dataset) else other)
        new.update(kw)
        return Dataset(new, self.coords)

    def drop_vars(self, names: Iterable[str]) -> "Dataset":
        names = set(names)
        return Dataset({k: v for k, v in self.data_vars.items() if k not in names}, self.coords)

    def isel(self, **indexers) -> "Dataset":
        coords = dict(self.coords)
        for dim, idx in indexers.items():
            if dim in coords:
                coords[dim] = coords[dim][idx]
        return Dataset({k: v.isel(**indexers) for k, v in self.data_vars.items()}, coords)

    def map(self, fn) -> "Dataset":
        return Dataset({k: fn(v) for k, v in self.data_vars.items()}, self.coords)

    def __repr__(self):
        body = ", ".join(f"{k}{v.dims}" for k, v in self.data_vars.items())
        return f"Dataset({body})"


def merge(datasets: Sequence[Dataset]) -> Dataset:
    """Union of variables (later datasets win), like xarray.merge for disjoint names."""
    out, coords = {}, {}
    for ds in datasets:
        out.update(ds.data_vars)
        coords.update(ds.coords)
    return Dataset(out, coords)


def concat_time(datasets: Sequence[Dataset]) -> Dataset:
    """Concatenate along 'time'; variables without a time axis are taken from the first."""
    first = datasets[0]
    out = {}
    for name, var in first.data_vars.items():
        if "time" in var.dims and all(name in d for d in datasets):
            ax = var.dims.index("time")
            out[name] = DataArray(np.concatenate([d[name].transpose(*var.dims).data for d in datasets], axis=ax), var.dims)
        else:
            out[name] = var
    coords = dict(first.coords)
    if all("time" in d.coords for d in datasets):
        coords["time"] = np.concatenate([np.atleast_1d(d.coords["time"]) for d in datasets])
    return Dataset(out, coords)


def from_xarray(ds) -> Dataset:
    """Convert a real xarray.Dataset (if the user has xarray)."""
    return Dataset({k: DataArray(np.asarray(v.data), v.dims) for k, v in ds.data_vars.items()},
                   {k: np.asarray(v.data) for k, v in ds.coords.items()})


def to_xarray(ds: Dataset):
    import xarray  # noqa: optional dependency
    return xarray.Dataset({k: (v.dims, v.data) for k, v in ds.data_vars.items()},
                          coords={k: v for k, v in ds.coords.items() if np.ndim(v) == 1})
