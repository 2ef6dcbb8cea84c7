"""Pytrees restricted to what the reference passes around: None, arrays, dicts, lists/tuples."""
import numpy as np


def tree_leaves(t):
    if t is None:
        return []
    if isinstance(t, dict):
        out = []
        for k in sorted(t):
            out += tree_leaves(t[k])
        return out
    if isinstance(t, (list, tuple)):
        out = []
        for v in t:
            out += tree_leaves(v)
        return out
    return [t]


def tree_map(fn, t, *rest):
    if t is None:
        return None
    if isinstance(t, dict):
        return {k: tree_map(fn, t[k], *[r[k] for r in rest]) for k in t}
    if isinstance(t, (list, tuple)):
        return type(t)(tree_map(fn, v, *[r[i] for r in rest]) for i, v in enumerate(t))
    return fn(t, *rest)
