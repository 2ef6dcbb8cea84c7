"""`normalization.InputsAndResiduals` of the reference (common/normalization.py) around the B200 GenCast.

Host path (`__call__`, `full_sampling`): the reference's Dataset arithmetic restated on xarray_lite -- inputs and
forcings are normalised with per-variable (optionally per-level) scales / locations, the wrapped predictor works in
normalised-residual space, predictions are un-normalised and the last input frame is added back for variables that are
also inputs (common/normalization.py:114-133, :147-161, :200-238).

Device path: `channel_transforms` flattens the same statistics into per-stacked-channel vectors, which
rollout.device_chunked_prediction feeds to gc_normalize_cast / gc_unnormalize_residual so that a whole autoregressive
rollout keeps its (physical-unit) input window on the GPU.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from .xarray_lite import DataArray, Dataset

_PRESERVED = ("batch", "lat", "lon")


def _stat_like(var: DataArray, stat: DataArray) -> np.ndarray:
    """Broadcasts a statistic (scalar or per-level) against a variable's dims."""
    s = np.asarray(stat.data, var.data.dtype)
    if s.ndim == 0:
        return s
    shape = [var.sizes[d] if d in stat.dims else 1 for d in var.dims]
    missing = [d for d in stat.dims if d not in var.dims]
    if missing:
        raise ValueError(f"normalisation statistic has dims {stat.dims} that the variable lacks ({var.dims})")
    order = [stat.dims.index(d) for d in var.dims if d in stat.dims]
    return np.transpose(s, order).reshape(shape)


def normalize(values: Dataset, scales: Dataset, locations: Optional[Dataset]) -> Dataset:
    """(x - location) / scale per variable (reference: common/normalization.py:31-50)."""
    out = {}
    for name, v in values.items():
        a = v.data
        if locations is not None and name in locations:
            a = a - _stat_like(v, locations[name])
        if name in scales:
            a = a / _stat_like(v, scales[name])
        out[name] = DataArray(a, v.dims)
    return Dataset(out, values.coords)


def unnormalize(values: Dataset, scales: Dataset, locations: Optional[Dataset]) -> Dataset:
    """x * scale + location per variable (reference: common/normalization.py:53-72)."""
    out = {}
    for name, v in values.items():
        a = v.data
        if name in scales:
            a = a * _stat_like(v, scales[name])
        if locations is not None and name in locations:
            a = a + _stat_like(v, locations[name])
        out[name] = DataArray(a, v.dims)
    return Dataset(out, values.coords)


def _channel_vector(template: Dataset, stats: Optional[Dataset], default: float) -> np.ndarray:
    """One value per stacked channel of `template` (sorted variables, non-(batch, lat, lon) dims row-major)."""
    parts = []
    for name in sorted(template.keys()):
        v = template[name]
        extra = [d for d in v.dims if d not in _PRESERVED]
        shape = [v.sizes[d] for d in extra]
        vals = np.full(shape or (), default, np.float32)
        if stats is not None and name in stats:
            st = stats[name]
            s = np.asarray(st.data, np.float32)
            if s.ndim:
                bshape = [v.sizes[d] if d in st.dims else 1 for d in extra]
                s = np.transpose(s, [st.dims.index(d) for d in extra if d in st.dims]).reshape(bshape)
            vals = np.broadcast_to(s, shape or ()).astype(np.float32)
        parts.append(np.asarray(vals, np.float32).reshape(-1))
    return np.concatenate(parts) if parts else np.zeros(0, np.float32)


class InputsAndResiduals:
    """Reference: common/normalization.py:75-238 (constructor :102-112)."""

    def __init__(self, predictor, stddev_by_level: Dataset, mean_by_level: Dataset, diffs_stddev_by_level: Dataset):
        self.predictor = predictor
        self._scales = stddev_by_level
        self._locations = mean_by_level
        self._residual_scales = diffs_stddev_by_level
        self._residual_locations = None

    # ---- host path
    def _unnormalize_prediction_and_add_input(self, inputs: Dataset, norm_prediction: Dataset) -> Dataset:
        out = {}
        for name, v in norm_prediction.items():
            if v.sizes.get("time") != 1:
                raise ValueError("normalization.InputsAndResiduals only supports predicting a single timestep.")
            one = Dataset({name: v}, norm_prediction.coords)
            if name in inputs:
                pred = unnormalize(one, self._residual_scales, self._residual_locations)[name]
                last = inputs[name].isel(time=slice(-1, None))
                out[name] = DataArray(pred.data + last.transpose(*pred.dims).data, pred.dims)
            else:
                out[name] = unnormalize(one, self._scales, self._locations)[name]
        return Dataset(out, norm_prediction.coords)

    def __call__(self, inputs: Dataset, targets_template: Dataset, forcings: Dataset, **kwargs) -> Dataset:
        norm_inputs = normalize(inputs, self._scales, self._locations)
        norm_forcings = normalize(forcings, self._scales, self._locations)
        return self._unnormalize_prediction_and_add_input(
            inputs, self.predictor(norm_inputs, targets_template, forcings=norm_forcings, **kwargs))

    def full_sampling(self, inputs: Dataset, targets_template: Dataset, forcings: Dataset, **kwargs) -> Dataset:
        norm_inputs = normalize(inputs, self._scales, self._locations)
        norm_forcings = normalize(forcings, self._scales, self._locations)
        # the template's values are never read by the sampler (only names / dims / coords): no transform needed
        norm_predictions = self.predictor.full_sampling(inputs=norm_inputs, targets_template=targets_template,
                                                        forcings=norm_forcings, **kwargs)
        return self._unnormalize_prediction_and_add_input(inputs, norm_predictions)

    # ---- device path
    def channel_transforms(self, inputs: Dataset, targets_template: Dataset, forcings: Dataset) -> Dict[str, np.ndarray]:
        """Per-stacked-channel vectors for gc_normalize_cast / gc_unnormalize_residual."""
        from .rollout import _channel_index
        t = dict(in_loc=_channel_vector(inputs, self._locations, 0.0), in_scale=_channel_vector(inputs, self._scales, 1.0),
                 frc_loc=_channel_vector(forcings, self._locations, 0.0), frc_scale=_channel_vector(forcings, self._scales, 1.0))
        # outputs: residual statistics for variables that are inputs, target statistics otherwise
        in_off, off = {}, 0
        for n in sorted(inputs.keys()):
            in_off[n] = off
            off += _channel_index(inputs[n])[1].size
        scale, loc, res_col = [], [], []
        for n in sorted(targets_template.keys()):
            one = Dataset({n: targets_template[n]}, targets_template.coords)
            extra, idx = _channel_index(targets_template[n])
            if n in inputs:
                scale.append(_channel_vector(one, self._residual_scales, 1.0))
                loc.append(np.zeros(idx.size, np.float32))
                iv = inputs[n]
                i_extra, i_idx = _channel_index(iv)
                last = np.take(i_idx, [iv.sizes["time"] - 1], axis=i_extra.index("time"))
                last = np.transpose(last, [i_extra.index(d) for d in extra])
                res_col.append(in_off[n] + last.reshape(-1))
            else:
                scale.append(_channel_vector(one, self._scales, 1.0))
                loc.append(_channel_vector(one, self._locations, 0.0))
                res_col.append(np.full(idx.size, -1, np.int64))
        t.update(out_scale=np.concatenate(scale), out_loc=np.concatenate(loc), res_col=np.concatenate(res_col).astype(np.int32))
        return t
