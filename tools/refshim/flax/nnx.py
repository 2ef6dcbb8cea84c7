"""numpy stand-in for the flax.nnx calls of the reference (Module, Linear, LayerNorm, Sequential,
Rngs, initializers, with_partitioning, remat).  Definitions restated from the Flax documentation:
Linear: y = x @ kernel + bias; LayerNorm: (x - mean) / sqrt(var + eps) [* scale + bias],
eps = 1e-6, fast variance var = E[x^2] - E[x]^2 clipped at 0."""
import numpy as np

DTYPE = np.float64


class Param:
    def __init__(self, value):
        self.value = np.asarray(value)


class Rngs:
    def __init__(self, default=0, **streams):
        self._rng = np.random.default_rng(default)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return lambda: int(self._rng.integers(0, 2 ** 31))

    def __call__(self):
        return int(self._rng.integers(0, 2 ** 31))


class Module:
    def named_params(self, prefix=""):
        """(path, Param) pairs, paths joined with '/', following attribute / dict-key / list-index names."""
        out = []

        def visit(obj, path):
            if isinstance(obj, Param):
                out.append((path, obj))
            elif isinstance(obj, Module):
                for k, v in vars(obj).items():
                    visit(v, f"{path}/{k}" if path else k)
            elif isinstance(obj, dict):
                for k, v in obj.items():
                    visit(v, f"{path}/{getattr(k, 'name', k)}")
            elif isinstance(obj, (list, tuple)):
                for i, v in enumerate(obj):
                    visit(v, f"{path}/{i}")

        visit(self, prefix)
        return out


class _Init:
    def __init__(self, kind, **kw):
        self.kind, self.kw = kind, kw

    def __call__(self, rng, shape):
        if self.kind == "zeros":
            return np.zeros(shape, DTYPE)
        if self.kind == "ones":
            return np.ones(shape, DTYPE)
        return rng.standard_normal(shape).astype(DTYPE) / np.sqrt(max(shape[0], 1))   # overwritten by the tests


class initializers:
    Initializer = _Init
    zeros_init = staticmethod(lambda: _Init("zeros"))
    ones_init = staticmethod(lambda: _Init("ones"))
    truncated_normal = staticmethod(lambda stddev=1.0: _Init("normal", stddev=stddev))
    xavier_uniform = staticmethod(lambda: _Init("normal"))
    variance_scaling = staticmethod(lambda *a, **k: _Init("normal"))


Initializer = _Init


def with_partitioning(init, spec=None, **k):
    return init


class Linear(Module):
    def __init__(self, in_features, out_features, *, use_bias=True, kernel_init=None, bias_init=None, rngs=None,
                 dtype=None, param_dtype=None, **kw):
        rng = np.random.default_rng(rngs.params() if rngs is not None else 0)
        self.kernel = Param((kernel_init or _Init("normal"))(rng, (in_features, out_features)))
        self.use_bias = use_bias
        if use_bias:
            self.bias = Param((bias_init or _Init("zeros"))(rng, (out_features,)))

    def __call__(self, x):
        y = np.asarray(x) @ self.kernel.value.astype(np.asarray(x).dtype)
        if self.use_bias:
            y = y + self.bias.value.astype(y.dtype)
        return y


class LayerNorm(Module):
    def __init__(self, num_features, *, use_scale=True, use_bias=True, epsilon=1e-6, feature_axes=-1,
                 scale_init=None, bias_init=None, rngs=None, use_fast_variance=True, **kw):
        self.epsilon = epsilon
        self.use_scale, self.use_bias = use_scale, use_bias
        if use_scale:
            self.scale = Param(np.ones(num_features, DTYPE))
        if use_bias:
            self.bias = Param(np.zeros(num_features, DTYPE))

    def __call__(self, x):
        mean = x.mean(-1, keepdims=True)
        var = np.maximum((x * x).mean(-1, keepdims=True) - mean * mean, 0.0)
        y = (x - mean) / np.sqrt(var + self.epsilon)
        if self.use_scale:
            y = y * self.scale.value
        if self.use_bias:
            y = y + self.bias.value
        return y


class Sequential(Module):
    def __init__(self, *layers):
        self.layers = list(layers)

    def __call__(self, x):
        for layer in self.layers:
            x = layer(x)
        return x


def remat(fn):
    return fn
