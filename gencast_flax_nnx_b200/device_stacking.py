"""Dataset <-> [grid node, channel] layout on the device.

Same mapping as stacking.py (which restates common/model_utils.py:594-725 and
gencast/denoiser.py:770-830 on the host), but the transposes run on the GPU: the variables'
raw arrays are packed back to back into one pinned staging buffer (plain memcpy), moved with a
single asynchronous H2D copy, and permuted into [lat*lon, channels] by strided device copies;
predictions take the reverse route with a single D2H copy.  The reference does the equivalent
inside its jitted function (xarray_jax + jnp transposes); doing it in numpy costs 10-20 ms per
12 h step at 2.5 deg, more than a fifth of the step.  torch is used for the copies and views only.
"""
from __future__ import annotations

import weakref
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .xarray_lite import DataArray, Dataset

PRESERVED = ("batch", "lat", "lon")
MAX_LEASED_OUTPUT_BUFFERS = 4


def pin_dataset(ds: Dataset) -> Dataset:
    """Copy of `ds` whose arrays live in page-locked host memory (what a loader that feeds the GPU
    would allocate).  DeviceStacker.to_nodes then copies each variable host -> device straight from
    the caller's arrays, without the packing memcpy through its own pinned staging buffer."""
    out = {}
    for name, v in ds.data_vars.items():
        a = np.ascontiguousarray(v.data, dtype=np.float32)
        t = torch.empty(a.shape, dtype=torch.float32, pin_memory=True)
        t.numpy()[...] = a
        out[name] = DataArray(t.numpy(), v.dims)
    return Dataset(out, ds.coords)


def _pinned_f32(a) -> Optional[torch.Tensor]:
    """torch view of a numpy array if it is contiguous fp32 in page-locked memory, else None."""
    if not isinstance(a, np.ndarray) or a.dtype != np.float32 or not a.flags.c_contiguous or a.size == 0:
        return None
    t = torch.from_numpy(a)
    return t if t.is_pinned() else None


class _Plan:
    """Where each variable sits in the packed host buffer and how it maps to channels."""

    def __init__(self, names: Sequence[str], ds: Dataset, sizes):
        self.entries = []          # (name, offset, shape, dims, channels)
        off = 0
        for n in names:
            v = ds[n]
            extra = [d for d in v.dims if d not in PRESERVED]
            c = int(np.prod([v.sizes[d] for d in extra], dtype=np.int64)) if extra else 1
            self.entries.append((n, off, tuple(v.shape), tuple(v.dims), c))
            off += int(np.prod(v.shape, dtype=np.int64))
        self.total = off
        self.channels = sum(e[4] for e in self.entries)
        self.signature = tuple((e[0], e[2], e[3]) for e in self.entries)


class DeviceStacker:
    """Packs Datasets of a fixed structure into [G, C] device tensors (batch element 0..B-1 separately)."""

    def __init__(self, device: torch.device):
        self.device = torch.device(device)
        self._plans = {}
        self._pinned = {}
        self._devbuf = {}
        self._out_pool = {}         # numel -> free pinned output buffers
        self._out_leased = 0

    def _plan(self, key: str, ds: Dataset, sizes) -> _Plan:
        names = sorted(ds.keys())
        plan = self._plans.get(key)
        sig = tuple((n, tuple(ds[n].shape), tuple(ds[n].dims)) for n in names)
        if plan is None or plan.signature != sig:
            plan = _Plan(names, ds, sizes)
            self._plans[key] = plan
            self._pinned[key] = torch.empty(max(plan.total, 1), dtype=torch.float32, pin_memory=True)
        return plan

    def to_nodes(self, key: str, ds: Dataset, sizes) -> torch.Tensor:
        """Dataset -> device tensor [lat*lon, batch, C] fp32 (sorted-name channel order, extra dims row-major)."""
        plan = self._plan(key, ds, sizes)
        B, n_lat, n_lon = sizes.get("batch", 1), sizes["lat"], sizes["lon"]
        if plan.total == 0:
            return torch.zeros(n_lat * n_lon, B, 0, dtype=torch.float32, device=self.device)
        pin = self._pinned[key]
        host = pin.numpy()
        dev = self._devbuf.get(key)
        if dev is None or dev.numel() != pin.numel():
            dev = torch.empty(pin.numel(), dtype=torch.float32, device=self.device)
            self._devbuf[key] = dev
        # Variables already in page-locked memory go host -> device directly; the rest are packed into
        # the stacker's pinned buffer first and moved with one copy per contiguous packed run.
        run_start = None
        for name, off, shape, dims, c in plan.entries:
            n = int(np.prod(shape, dtype=np.int64))
            direct = _pinned_f32(ds[name].data)
            if direct is not None:
                if run_start is not None:
                    dev[run_start:off].copy_(pin[run_start:off], non_blocking=True)
                    run_start = None
                dev[off:off + n].copy_(direct.view(-1), non_blocking=True)
            else:
                host[off:off + n] = np.asarray(ds[name].data, dtype=np.float32).reshape(-1)
                if run_start is None:
                    run_start = off
        if run_start is not None:
            dev[run_start:plan.total].copy_(pin[run_start:plan.total], non_blocking=True)
        blocks = []
        for name, off, shape, dims, c in plan.entries:
            n = int(np.prod(shape, dtype=np.int64))
            t = dev[off:off + n].view(shape)
            extra = [d for d in dims if d not in PRESERVED]
            have = [d for d in PRESERVED if d in dims]
            t = t.permute([dims.index(d) for d in have + extra])
            t = t.reshape(t.shape[:len(have)] + (c,))
            # insert missing preserved dims, broadcast, then (batch, lat, lon, c) -> (lat, lon, batch, c)
            full, k = [], 0
            for d in PRESERVED:
                if d in have:
                    full.append(t.shape[k]); k += 1
                else:
                    full.append(1)
            t = t.reshape(full + [c]).expand(B, n_lat, n_lon, c).permute(1, 2, 0, 3)
            blocks.append(t.reshape(n_lat * n_lon, B, c))
        return torch.cat(blocks, dim=-1).contiguous()

    def from_nodes(self, nodes: torch.Tensor, template: Dataset) -> Dataset:
        """Device [lat*lon, batch, C] -> Dataset shaped like `template` (one D2H copy)."""
        sizes = template.sizes
        n_lat, n_lon = sizes["lat"], sizes["lon"]
        B = nodes.shape[1]
        names = sorted(template.keys())
        pieces, meta, i = [], [], 0
        for name in names:
            tv = template[name]
            if not all(d in tv.dims for d in PRESERVED):
                raise ValueError(f"stacked_to_dataset requires all variables to have {PRESERVED} dimensions, "
                                 f"but found only {tv.dims}.")
            extra = [d for d in tv.dims if d not in PRESERVED]
            eshape = [tv.sizes[d] for d in extra]
            c = int(np.prod(eshape, dtype=np.int64)) if extra else 1
            blk = nodes[:, :, i:i + c].reshape([n_lat, n_lon, B] + eshape)
            i += c
            cur = ["lat", "lon", "batch"] + extra
            blk = blk.permute([cur.index(d) for d in tv.dims]).contiguous()
            pieces.append(blk.reshape(-1))
            meta.append((name, tuple(blk.shape), tuple(tv.dims)))
        if i != nodes.shape[-1]:
            raise ValueError(f"Expected {i} channels but found {nodes.shape[-1]}")
        flat = torch.cat(pieces)
        pin, leased = self._lease_output(flat.numel())
        pin.copy_(flat, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        host = pin.numpy()
        if leased:
            # The returned arrays are views of this page-locked buffer (no second host copy).  The buffer
            # goes back to the pool when the last view of it is garbage collected; a caller that keeps
            # results alive simply keeps their buffers, and past MAX_LEASED_OUTPUT_BUFFERS outstanding
            # results the arrays are copied into ordinary memory instead.
            weakref.finalize(host, self._release_output, pin)
        out, off = {}, 0
        for name, shape, dims in meta:
            n = int(np.prod(shape, dtype=np.int64))
            view = host[off:off + n].reshape(shape)
            out[name] = DataArray(view if leased else view.copy(), dims)
            off += n
        return Dataset(out, template.coords)

    def _lease_output(self, numel: int):
        free = self._out_pool.setdefault(numel, [])
        if free:
            self._out_leased += 1
            return free.pop(), True
        if self._out_leased < MAX_LEASED_OUTPUT_BUFFERS:
            self._out_leased += 1
            return torch.empty(numel, dtype=torch.float32, pin_memory=True), True
        key = ("out_copy", numel)
        pin = self._pinned.get(key)
        if pin is None:
            pin = torch.empty(numel, dtype=torch.float32, pin_memory=True)
            self._pinned[key] = pin
        return pin, False

    def _release_output(self, pin: torch.Tensor) -> None:
        self._out_leased -= 1
        self._out_pool.setdefault(pin.numel(), []).append(pin)
