import numpy as np


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def swish(x):
    return x * sigmoid(x)


silu = swish


def relu(x):
    return np.maximum(x, 0)


def gelu(x, approximate=True):
    """jax.nn.gelu: tanh approximation by default."""
    if approximate:
        return 0.5 * x * (1.0 + np.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * x ** 3)))
    from scipy.special import erf
    return 0.5 * x * (1.0 + erf(x / np.sqrt(2.0)))


def softmax(x, axis=-1):
    m = np.max(x, axis=axis, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=axis, keepdims=True)
