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

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from .xarray_lite import DataArray, Dataset

PRESERVED = ("batch", "lat", "lon")


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
        for name, off, shape, dims, c in plan.entries:
            n = int(np.prod(shape, dtype=np.int64))
            host[off:off + n] = np.asarray(ds[name].data, dtype=np.float32).reshape(-1)
        dev = pin.to(self.device, non_blocking=True)
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
        key = ("out", flat.numel())
        pin = self._pinned.get(key)
        if pin is None:
            pin = torch.empty(flat.numel(), dtype=torch.float32, pin_memory=True)
            self._pinned[key] = pin
        pin.copy_(flat, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        host = pin.numpy()
        out, off = {}, 0
        for name, shape, dims in meta:
            n = int(np.prod(shape, dtype=np.int64))
            out[name] = DataArray(host[off:off + n].reshape(shape).copy(), dims)
            off += n
        return Dataset(out, template.coords)
