"""numpy stand-in for the parts of `jax` the reference's hot-path modules touch."""
import functools
import numpy as _np
from . import numpy  # noqa: F401  (jax.numpy)
from . import nn, lax, tree_util, sharding, experimental, random  # noqa: F401

Array = _np.ndarray


def vmap(*a, **k):
    raise NotImplementedError("refshim: jax.vmap is only used by attention types the default config does not select")


def custom_vjp(fn=None, nondiff_argnums=()):
    if fn is None:
        return functools.partial(custom_vjp, nondiff_argnums=nondiff_argnums)
    fn.defvjp = lambda *a, **k: None
    return fn


def jit(fn=None, **k):
    return fn if fn is not None else (lambda f: f)


class _Typing:
    ArrayLike = object
    DTypeLike = object


typing = _Typing()


class _Tree:
    @staticmethod
    def map(fn, *trees):
        from .tree_util import tree_map
        return tree_map(fn, *trees)


tree = _Tree()
