"""jraph.segment_sum restated: unsorted scatter-add, zeros for empty segments (= jax.ops.segment_sum)."""
import numpy as np
from typing import Any, Callable

NodeFeatures = Any
ArrayTree = Any
AggregateEdgesToNodesFn = Callable
AggregateNodesToGlobalsFn = Callable
AggregateEdgesToGlobalsFn = Callable


def segment_sum(data, segment_ids, num_segments=None, indices_are_sorted=False, unique_indices=False):
    data = np.asarray(data)
    out = np.zeros((num_segments,) + data.shape[1:], dtype=data.dtype)
    np.add.at(out, np.asarray(segment_ids), data)
    return out
