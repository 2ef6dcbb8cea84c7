"""Spherical white noise (gencast/samplers_utils.py:250-346).  CPU: the float64 host restatement
(oracle/spherical_noise_oracle.py) has the statistical contract of the reference's generator -- zero mean, unit variance
at every latitude including the poles, single-valued at the poles (isotropy).  GPU: gc_sh_synthesis reproduces the
restatement on the same coefficients and in the sampler's state layout."""
import numpy as np
import pytest
import torch

from gencast_flax_nnx_b200 import graph
from gencast_flax_nnx_b200.spherical_noise import amplitude_table, legendre_table


def test_legendre_orthonormality():
    # Gauss-Legendre quadrature: (1/2) int Pbar_lm Pbar_l'm dt = delta_ll' * (1 for m = 0, 2 for m > 0) / ... (4 pi norm)
    L = 24
    x, w = np.polynomial.legendre.leggauss(64)
    P = legendre_table(L, x)
    for m in (0, 1, 5, 17):
        gram = np.einsum("lj,kj,j->lk", P[m, m:], P[m, m:], w) / 2.0
        expect = np.eye(L - m) * (1.0 if m == 0 else 2.0)
        np.testing.assert_allclose(gram, expect, atol=1e-10)


def test_noise_statistics_of_the_host_restatement():
    from oracle import spherical_noise_oracle as so
    lat, lon = graph.regular_grid(5.0)
    table = amplitude_table(lat, len(lon))
    L = table.shape[0]
    rng = np.random.default_rng(0)
    f = so.synthesize(rng.standard_normal((2, L, 4000, L)), table, len(lon))
    assert abs(f.mean()) < 5e-3
    var = f.var(axis=(0, 2))                               # per latitude
    np.testing.assert_allclose(var, 1.0, atol=0.06)
    # poles: one physical point -> identical value at every longitude
    assert np.abs(f[:, 0, :] - f[:, 0, :1]).max() < 1e-9 and np.abs(f[:, -1, :] - f[:, -1, :1]).max() < 1e-9
    # neighbouring longitudes are strongly correlated near the poles and weakly at the equator
    eq = len(lat) // 2
    corr = lambda a, b: float(np.mean(a * b) / np.sqrt(np.mean(a * a) * np.mean(b * b)))
    assert corr(f[:, 1, 0], f[:, 1, 1]) > 0.95
    assert abs(corr(f[:, eq, 0], f[:, eq, 1])) < 0.75


def test_product_generator_has_no_cpu_path():
    from gencast_flax_nnx_b200.spherical_noise import SphericalNoise
    lat, lon = graph.regular_grid(30.0)
    with pytest.raises(RuntimeError):
        SphericalNoise(lat, lon, device="cpu")


@pytest.mark.gpu
@pytest.mark.parametrize("res,members,channels", [(30.0, 1, 5), (10.0, 3, 82), (2.5, 2, 82)])
def test_gpu_synthesis_matches_host_restatement(cuda_device, res, members, channels):
    from oracle import spherical_noise_oracle as so
    from gencast_flax_nnx_b200.spherical_noise import SphericalNoise
    lat, lon = graph.regular_grid(res)
    sn = SphericalNoise(lat, lon, cuda_device)
    g = torch.Generator(device=cuda_device).manual_seed(7)
    coef = sn.draw_coefficients(members * channels, g)
    got = sn.synthesize(coef, channels, members).cpu().numpy()
    ref = so.synthesize(coef.cpu().numpy(), amplitude_table(lat, len(lon)), len(lon))       # [F, n_lat, n_lon]
    ref = ref.reshape(members, channels, len(lat) * len(lon)).transpose(0, 2, 1).reshape(-1, channels)
    assert np.abs(got - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())
    # unit marginal variance (loose: few fields) and single-valued poles
    assert 0.8 < got.var() < 1.2
    pole = got.reshape(members, len(lat), len(lon), channels)[:, 0]
    assert np.abs(pole - pole[:, :1]).max() < 1e-4
    again = sn.synthesize(coef, channels, members).cpu().numpy()
    assert np.array_equal(got, again)
