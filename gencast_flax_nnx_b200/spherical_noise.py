"""Isotropic white noise on the sphere for the sampler's initial state.

The reference draws its noise in spherical-harmonic space (gencast/samplers_utils.py:250-346):
total wavenumbers l = 0 .. n_lon/2 - 1 carry equal power 1/(n_lon/2), split evenly over the 2l+1
real harmonics of each l, and the field is synthesised on the lat/lon grid with dinosaur's real
spherical harmonics, so that every grid point has unit marginal variance and the field is
rotation invariant (in particular single-valued at the poles, unlike white noise per grid cell).

This module restates that construction: geodesy-normalised associated Legendre functions by the
standard stable recurrence (float64, host, once), then per draw a Legendre synthesis (one batched
matmul per zonal wavenumber m) and an inverse real FFT along longitude.  It is set-up work of a
sampling step, not part of the 40-evaluation hot loop; the two transforms are torch library calls
(cuBLAS / cuFFT), not hand-written kernels.  dinosaur is not installable here, so agreement with
its exact coefficient ordering / random stream is unpinned; the statistical contract (zero mean, unit
variance at every latitude, flat spectrum, isotropy) is tested.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch


def legendre_table(n_wavenumbers: int, sin_lat: np.ndarray) -> np.ndarray:
    """Pbar[m, l, j]: 4-pi-normalised associated Legendre functions (sqrt(2) included for m > 0),
    zero for l < m.  (1 / 4 pi) * integral of (Pbar_lm cos(m phi))^2 over the sphere = 1."""
    L = n_wavenumbers
    t = np.asarray(sin_lat, np.float64)
    u = np.sqrt(np.maximum(1.0 - t * t, 0.0))
    out = np.zeros((L, L, t.shape[0]), np.float64)
    pmm = np.ones_like(t)
    for m in range(L):
        if m == 1:
            pmm = math.sqrt(3.0) * u
        elif m > 1:
            pmm = u * math.sqrt((2.0 * m + 1.0) / (2.0 * m)) * pmm
        out[m, m] = pmm
        if m + 1 < L:
            out[m, m + 1] = math.sqrt(2.0 * m + 3.0) * t * pmm
        for l in range(m + 2, L):
            a = math.sqrt((4.0 * l * l - 1.0) / (l * l - m * m))
            b = math.sqrt(((l - 1.0) ** 2 - m * m) / (4.0 * (l - 1.0) ** 2 - 1.0))
            out[m, l] = a * (t * out[m, l - 1] - b * out[m, l - 2])
    return out


class SphericalNoise:
    """Unit-variance isotropic white noise fields on an equiangular lat/lon grid (poles included)."""

    def __init__(self, grid_lat, grid_lon, device=None):
        lat = np.asarray(grid_lat, np.float64)
        self.n_lat, self.n_lon = len(lat), len(grid_lon)
        self.L = max(1, self.n_lon // 2)                     # gencast/samplers_utils.py:336
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        table = legendre_table(self.L, np.sin(np.deg2rad(lat)))
        # per-l amplitude: power 1/L per total wavenumber over 2l+1 harmonics (samplers_utils.py:316-322)
        amp = np.sqrt((1.0 / self.L) / (2.0 * np.arange(self.L) + 1.0))
        self.table = torch.from_numpy((table * amp[None, :, None]).astype(np.float32)).to(self.device)   # [m, l, lat]

    def sample(self, n_fields: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """[n_fields, n_lat, n_lon] fp32 on the device."""
        L, n = self.L, self.n_lon
        coef = torch.randn(2, L, n_fields, L, generator=generator, device=self.device)     # (cos|sin, m, field, l)
        # Legendre synthesis per zonal wavenumber: [m, field, l] @ [m, l, lat] -> [m, field, lat]
        a = torch.bmm(coef[0], self.table)
        b = torch.bmm(coef[1], self.table)
        spec = torch.zeros(n_fields, self.n_lat, n // 2 + 1, dtype=torch.complex64, device=self.device)
        spec[:, :, :L] = torch.complex(a, -b).permute(1, 2, 0) * (0.5 * n)
        spec[:, :, 0] = torch.complex(a[0], torch.zeros_like(a[0])) * float(n)
        return torch.fft.irfft(spec, n=n, dim=-1)

    def sample_nodes(self, channels: int, members: int = 1, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """[members * n_lat * n_lon, channels]: the sampler state layout (node index = lat * n_lon + lon)."""
        f = self.sample(members * channels, generator).reshape(members, channels, self.n_lat * self.n_lon)
        return f.permute(0, 2, 1).reshape(members * self.n_lat * self.n_lon, channels).contiguous()
