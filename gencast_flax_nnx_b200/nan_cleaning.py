"""`nan_cleaning.NaNCleaner` of the reference (gencast/nan_cleaning.py:24-160): the wrapped predictor sees inputs and
forcings with the NaNs of one variable (sea-surface temperature over land) replaced by a fill value; optionally the NaNs
are put back into the predictions.  Host path on xarray_lite; `channel_fill` gives the per-stacked-channel vectors the
device-resident rollout uses (gc_normalize_cast / gc_unnormalize_residual)."""
from __future__ import annotations

from typing import Optional

import numpy as np

from .xarray_lite import DataArray, Dataset


class NaNCleaner:
    def __init__(self, predictor, var_to_clean: str, fill_value: Dataset, reintroduce_nans: bool = False):
        self.predictor = predictor
        self._fill_value = fill_value[var_to_clean]
        self._var_to_clean = var_to_clean
        self._reintroduce_nans = reintroduce_nans

    def _clean(self, dataset: Dataset) -> Dataset:                                   # nan_cleaning.py:47-53
        v = dataset[self._var_to_clean]
        fill = np.asarray(self._fill_value.data, v.data.dtype)
        return dataset.assign({self._var_to_clean: DataArray(np.where(np.isnan(v.data), fill, v.data), v.dims)})

    def _maybe_reintroduce_nans(self, stale_inputs: Dataset, predictions: Dataset) -> Dataset:   # :55-64
        if self._var_to_clean in predictions.keys():
            iv = stale_inputs[self._var_to_clean]
            mask = np.isnan(iv.data).any(axis=iv.dims.index("time"), keepdims=True)
            pv = predictions[self._var_to_clean]
            mask = DataArray(mask, iv.dims).transpose(*pv.dims).data
            predictions = predictions.assign({self._var_to_clean: DataArray(np.where(mask, np.nan, pv.data).astype(pv.data.dtype), pv.dims)})
        return predictions

    def _run(self, fn, inputs: Dataset, targets_template: Dataset, forcings: Optional[Dataset], **kwargs) -> Dataset:
        original_inputs = inputs
        if self._var_to_clean in inputs.keys():
            inputs = self._clean(inputs)
        if forcings is not None and self._var_to_clean in forcings.keys():
            forcings = self._clean(forcings)
        predictions = fn(inputs, targets_template, forcings, **kwargs)
        if self._reintroduce_nans:
            predictions = self._maybe_reintroduce_nans(original_inputs, predictions)
        return predictions

    def __call__(self, inputs, targets_template, forcings=None, **kwargs):             # :66-86
        return self._run(self.predictor, inputs, targets_template, forcings, **kwargs)

    def full_sampling(self, inputs, targets_template, forcings=None, **kwargs):        # :129-156
        return self._run(lambda i, t, f, **kw: self.predictor.full_sampling(inputs=i, targets_template=t, forcings=f, **kw),
                         inputs, targets_template, forcings, **kwargs)

    def channel_fill(self, ds: Dataset) -> np.ndarray:
        """Per-stacked-channel fill values of `ds` (NaN = leave the channel alone)."""
        from .normalization import _channel_vector
        fill = Dataset({self._var_to_clean: self._fill_value}, {}) if self._var_to_clean in ds else None
        return _channel_vector(ds, fill, np.nan)
