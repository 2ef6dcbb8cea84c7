"""Minimal numpy-backed stand-in for the xarray calls the reference's sampler, preconditioner and stacking code
make (gencast/dpm_solver_plus_plus_2s.py, common/model_utils.py:145-167, :594-725).  TEST INFRASTRUCTURE ONLY.

Semantics restated from the xarray documentation, for exactly the calls used:
  * arithmetic broadcasts by dimension NAME; result dims = dims of the left operand followed by the right
    operand's new dims (xarray's `broadcast` ordering: order of first appearance);
  * Variable.stack(channels=[d1, d2, ...]) moves the listed dims to the end and flattens them row-major;
    unstack is its inverse; set_dims inserts missing dims (broadcast) and orders as requested;
  * Dataset.to_array() broadcasts every variable to the union of dims (order of first appearance over the
    variables) and stacks them along a new leading 'variable' dim; DataArray.to_dataset(dim='variable') undoes it;
  * Dataset.assign / drop_vars / map / item access as in xarray.
Coordinates are carried as plain name -> 1-D numpy arrays attached to Datasets / DataArrays (no index alignment:
every operand in the exercised code paths shares its coordinates).
"""
from __future__ import annotations

import numpy as np


def _union_dims(*dim_lists):
    out = []
    for dims in dim_lists:
        for d in dims:
            if d not in out:
                out.append(d)
    return tuple(out)


def _expand(data, dims, out_dims):
    """View of `data` (dims) broadcastable against arrays with dims `out_dims` (missing dims -> size 1)."""
    perm = [dims.index(d) for d in out_dims if d in dims]
    data = np.transpose(data, perm)
    shape = []
    it = iter(data.shape)
    for d in out_dims:
        shape.append(next(it) if d in dims else 1)
    return data.reshape(shape)


class Variable:
    def __init__(self, dims, data):
        if isinstance(dims, str):
            dims = (dims,)
        self.dims = tuple(dims)
        self.data = np.asarray(data)
        if self.data.ndim != len(self.dims):
            raise ValueError(f"{self.data.shape} vs {self.dims}")

    # ---- bookkeeping
    @property
    def sizes(self):
        return dict(zip(self.dims, self.data.shape))

    @property
    def shape(self):
        return self.data.shape

    @property
    def dtype(self):
        return self.data.dtype

    @property
    def size(self):
        return self.data.size

    @property
    def variable(self):
        return self

    def astype(self, dt):
        return type(self)._like(self, self.dims, self.data.astype(dt))

    @classmethod
    def _like(cls, src, dims, data):
        return Variable(dims, data)

    # ---- reshaping
    def transpose(self, *dims):
        if Ellipsis in dims:
            i = dims.index(Ellipsis)
            named = [d for d in dims if d is not Ellipsis]
            rest = [d for d in self.dims if d not in named]
            dims = tuple(dims[:i]) + tuple(rest) + tuple(dims[i + 1:])
        perm = [self.dims.index(d) for d in dims]
        return type(self)._like(self, dims, np.transpose(self.data, perm))

    def stack(self, **kw):
        (new, old), = kw.items()
        keep = [d for d in self.dims if d not in old]
        v = self.transpose(*keep, *old)
        n = int(np.prod([self.sizes[d] for d in old]))
        return type(self)._like(self, tuple(keep) + (new,), v.data.reshape(v.data.shape[:len(keep)] + (n,)))

    def unstack(self, mapping):
        (old, sizes), = mapping.items()
        ax = self.dims.index(old)
        if ax != len(self.dims) - 1:
            raise NotImplementedError("refshim: unstack of a non-trailing dim")
        new_dims = tuple(sizes.keys())
        shape = self.data.shape[:ax] + tuple(int(s) for s in sizes.values())
        return type(self)._like(self, self.dims[:ax] + new_dims, self.data.reshape(shape))

    def set_dims(self, dims):
        out_dims = tuple(dims.keys())
        extra = [d for d in self.dims if d not in out_dims]
        if extra:
            raise ValueError(f"set_dims would drop {extra}")
        data = _expand(self.data, list(self.dims), out_dims)
        data = np.broadcast_to(data, tuple(int(dims[d]) for d in out_dims))
        return type(self)._like(self, out_dims, data)

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        data, dims = self.data, list(self.dims)
        for dim, idx in indexers.items():
            ax = dims.index(dim)
            if isinstance(idx, slice):
                data = data[(slice(None),) * ax + (idx,)]
            else:
                data = np.take(data, idx, axis=ax)
                dims.pop(ax)
        return type(self)._like(self, tuple(dims), data)

    @staticmethod
    def concat(variables, dim):
        first = variables[0]
        ax = first.dims.index(dim)
        vs = [v.transpose(*first.dims).data for v in variables]
        return Variable(first.dims, np.concatenate(vs, axis=ax))

    # ---- arithmetic by dimension name
    def _binary(self, other, op, reflexive=False):
        if isinstance(other, Dataset):
            return NotImplemented
        if isinstance(other, Variable):
            dims = _union_dims(self.dims, other.dims)
            a, b = _expand(self.data, list(self.dims), dims), _expand(other.data, list(other.dims), dims)
        else:
            dims, a, b = self.dims, self.data, other
        res = op(b, a) if reflexive else op(a, b)
        return type(self)._like(self, dims, res)

    def __add__(self, o): return self._binary(o, np.add)
    def __radd__(self, o): return self._binary(o, np.add, True)
    def __sub__(self, o): return self._binary(o, np.subtract)
    def __rsub__(self, o): return self._binary(o, np.subtract, True)
    def __mul__(self, o): return self._binary(o, np.multiply)
    def __rmul__(self, o): return self._binary(o, np.multiply, True)
    def __truediv__(self, o): return self._binary(o, np.divide)
    def __rtruediv__(self, o): return self._binary(o, np.divide, True)
    def __pow__(self, o): return self._binary(o, np.power)
    def __neg__(self): return type(self)._like(self, self.dims, -self.data)


class DataArray(Variable):
    def __init__(self, data=None, coords=None, dims=None, name=None):
        if isinstance(data, Variable):
            dims = data.dims if dims is None else dims
            data = data.data
        data = np.asarray(data)
        if dims is None:
            raise ValueError("refshim DataArray needs dims")
        super().__init__(dims, data)
        self.coords = {k: (v if isinstance(v, DataArray) else DataArray(np.asarray(v), dims=(k,)))
                       for k, v in (coords or {}).items() if np.ndim(getattr(v, "data", v)) == 1}
        self.name = name

    @classmethod
    def _like(cls, src, dims, data):
        coords = {k: v for k, v in getattr(src, "coords", {}).items() if k in dims}
        return DataArray(data, coords=coords, dims=dims, name=getattr(src, "name", None))

    @property
    def variable(self):
        return Variable(self.dims, self.data)

    def __getattr__(self, name):
        coords = self.__dict__.get("coords", {})
        if name in coords:
            return coords[name]
        raise AttributeError(name)

    def drop(self, name):
        return DataArray(self.data, coords={k: v for k, v in self.coords.items() if k != name}, dims=self.dims, name=self.name)

    drop_vars = drop

    def to_dataset(self, dim):
        ax = self.dims.index(dim)
        names = [str(n) for n in self.coords[dim].data]
        rest = tuple(d for d in self.dims if d != dim)
        coords = {k: v for k, v in self.coords.items() if k != dim}
        return Dataset({n: DataArray(np.take(self.data, i, axis=ax), coords=coords, dims=rest, name=n)
                        for i, n in enumerate(names)}, coords=coords)


class Dataset:
    def __init__(self, data_vars=None, coords=None):
        self._vars = {}
        self.coords = {}
        for k, v in (coords or {}).items():
            self.coords[k] = v if isinstance(v, DataArray) else DataArray(np.asarray(v), dims=(k,))
        for k, v in (data_vars or {}).items():
            self[k] = v

    # ---- mapping
    def __setitem__(self, key, value):
        if isinstance(value, tuple):
            value = DataArray(value[1], dims=value[0])
        if not isinstance(value, DataArray):
            value = DataArray(value.data, dims=value.dims)
        for k, c in value.coords.items():
            self.coords.setdefault(k, c)
        self._vars[key] = value

    def __getitem__(self, key):
        if isinstance(key, (list, tuple)):
            return Dataset({k: self._vars[k] for k in key}, self.coords)
        v = self._vars[key]
        return DataArray(v.data, coords={k: c for k, c in self.coords.items() if k in v.dims}, dims=v.dims, name=key)

    def __contains__(self, key):
        return key in self._vars

    def __iter__(self):
        return iter(self._vars)

    def keys(self):
        return self._vars.keys()

    def items(self):
        return [(k, self[k]) for k in self._vars]

    def __len__(self):
        return len(self._vars)

    @property
    def data_vars(self):
        return {k: self[k] for k in self._vars}

    @property
    def variables(self):
        out = {k: v.variable for k, v in self._vars.items()}
        out.update({k: v.variable for k, v in self.coords.items()})
        return out

    @property
    def sizes(self):
        out = {}
        for v in self._vars.values():
            out.update(v.sizes)
        return out

    dims = sizes

    def assign(self, other=None, **kw):
        new = Dataset(dict(self._vars), self.coords)
        for src in ((other._vars if isinstance(other, Dataset) else other) or {}, kw):
            for k, v in src.items():
                new[k] = v
        return new

    def drop_vars(self, names):
        names = set([names] if isinstance(names, str) else names)
        return Dataset({k: v for k, v in self._vars.items() if k not in names}, self.coords)

    def map(self, fn):
        return Dataset({k: fn(self[k]) for k in self._vars}, self.coords)

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        return Dataset({k: v.isel({d: i for d, i in indexers.items() if d in v.dims}) for k, v in self._vars.items()},
                       {k: (c.isel({k: indexers[k]}) if k in indexers and isinstance(indexers[k], slice) else c)
                        for k, c in self.coords.items() if not (k in indexers and not isinstance(indexers[k], slice))})

    def to_array(self, dim="variable"):
        names = list(self._vars)
        dims = _union_dims(*[self._vars[n].dims for n in names])
        sizes = self.sizes
        shape = tuple(sizes[d] for d in dims)
        stacked = np.stack([np.broadcast_to(_expand(self._vars[n].data, list(self._vars[n].dims), dims), shape)
                            for n in names])
        coords = {k: v for k, v in self.coords.items() if k in dims}
        coords[dim] = DataArray(np.asarray(names, dtype=object), dims=(dim,))
        return DataArray(stacked, coords=coords, dims=(dim,) + dims)

    # ---- arithmetic: applied per variable
    def _binary(self, other, op, reflexive=False):
        out = {}
        for k in self._vars:
            o = other[k] if isinstance(other, Dataset) else other
            a = self[k]
            out[k] = a._binary(o, op, reflexive)
        return Dataset(out, self.coords)

    def __add__(self, o): return self._binary(o, np.add)
    def __radd__(self, o): return self._binary(o, np.add, True)
    def __sub__(self, o): return self._binary(o, np.subtract)
    def __mul__(self, o): return self._binary(o, np.multiply)
    def __rmul__(self, o): return self._binary(o, np.multiply, True)
    def __truediv__(self, o): return self._binary(o, np.divide)


def concat(objs, dim):
    first = objs[0]
    v = Variable.concat([o.variable for o in objs], dim)
    return DataArray(v.data, coords=getattr(first, "coords", None), dims=v.dims)
