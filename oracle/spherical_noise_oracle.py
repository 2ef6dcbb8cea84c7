"""Float64 host restatement of the reference's spherical-harmonic noise synthesis.  TEST INFRASTRUCTURE ONLY.

Reference: gencast/samplers_utils.py:250-346 (`sample`: coefficients per (total wavenumber, longitude wavenumber),
triangular mask |m| <= l, per-l normalisation sqrt(power_l / (2 l + 1)) * sqrt(4 pi), inverse transform by
dinosaur.spherical_harmonic.RealSphericalHarmonics.to_nodal).  dinosaur is a third-party dependency that is absent from
/root/reference and not installable here (requirements.txt:17, unpinned), so its published algorithm is restated: real
spherical harmonics orthonormal on the unit sphere, Y_lm = Pbar_lm(sin lat) {cos, sin}(m lon) / sqrt(4 pi) with the
4-pi-normalised associated Legendre functions; the reference's sqrt(4 pi) factor cancels that normalisation.
PARITY UNPINNED with respect to dinosaur's coefficient ordering and JAX's random stream: the checkable contract is
(a) the synthesis of given coefficients and (b) the statistics of the resulting fields.
"""
from __future__ import annotations

import numpy as np


def synthesize(coef: np.ndarray, table: np.ndarray, n_lon: int) -> np.ndarray:
    """coef [2, L, F, L] (cos | sin, m, field, l), table [L (m), L (l), n_lat] float64 -> fields [F, n_lat, n_lon]."""
    coef = np.asarray(coef, np.float64)
    table = np.asarray(table, np.float64)
    L = table.shape[0]
    mask = (np.arange(L)[:, None] <= np.arange(L)[None, :]).astype(np.float64)          # m <= l  (samplers_utils.py:305-311)
    a = np.einsum("mfl,ml,mlj->fjm", coef[0], mask, table)                               # [F, n_lat, m]
    b = np.einsum("mfl,ml,mlj->fjm", coef[1], mask, table)
    b[:, :, 0] = 0.0                                                                     # sin(0) harmonic does not exist
    phi = 2.0 * np.pi * np.arange(n_lon) / n_lon
    ang = np.arange(L)[:, None] * phi[None, :]
    return a @ np.cos(ang) + b @ np.sin(ang)
