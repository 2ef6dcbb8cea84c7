"""Statistical contract of the spherical white noise (gencast/samplers_utils.py:250-346): zero mean,
unit variance at every latitude including the poles, single-valued at the poles (isotropy), flat
power over total wavenumbers."""
import numpy as np
import torch

from gencast_flax_nnx_b200 import graph
from gencast_flax_nnx_b200.spherical_noise import SphericalNoise, legendre_table


def test_legendre_orthonormality():
    # Gauss-Legendre quadrature: (1/2) int Pbar_lm Pbar_l'm dt = delta_ll' * (1 for m = 0, 2 for m > 0) / ... (4 pi norm)
    L = 24
    x, w = np.polynomial.legendre.leggauss(64)
    P = legendre_table(L, x)
    for m in (0, 1, 5, 17):
        gram = np.einsum("lj,kj,j->lk", P[m, m:], P[m, m:], w) / 2.0
        expect = np.eye(L - m) * (1.0 if m == 0 else 2.0)
        np.testing.assert_allclose(gram, expect, atol=1e-10)


def test_noise_statistics():
    lat, lon = graph.regular_grid(5.0)
    sn = SphericalNoise(lat, lon)
    g = torch.Generator().manual_seed(0)
    f = sn.sample(4000, g).numpy()
    assert abs(f.mean()) < 5e-3
    var = f.var(axis=(0, 2))                               # per latitude
    np.testing.assert_allclose(var, 1.0, atol=0.06)
    # poles: one physical point -> identical value at every longitude
    assert np.abs(f[:, 0, :] - f[:, 0, :1]).max() < 1e-4 and np.abs(f[:, -1, :] - f[:, -1, :1]).max() < 1e-4
    # neighbouring longitudes are strongly correlated near the poles and weakly at the equator
    eq = len(lat) // 2
    corr = lambda a, b: float(np.mean(a * b) / np.sqrt(np.mean(a * a) * np.mean(b * b)))
    assert corr(f[:, 1, 0], f[:, 1, 1]) > 0.95
    assert abs(corr(f[:, eq, 0], f[:, eq, 1])) < 0.75
    nodes = sn.sample_nodes(82, members=2, generator=g)
    assert nodes.shape == (2 * len(lat) * len(lon), 82)
