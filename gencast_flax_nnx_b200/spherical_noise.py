"""Isotropic white noise on the sphere for the sampler's initial state and its stochastic churn
(SURVEY.md §8f item 1).

The reference draws its noise in spherical-harmonic space (gencast/samplers_utils.py:250-346):
total wavenumbers l = 0 .. n_lon/2 - 1 carry equal power 1/(n_lon/2), split evenly over the 2l+1
real harmonics of each l, and the field is synthesised on the lat/lon grid with dinosaur's real
spherical harmonics, so that every grid point has unit marginal variance and the field is
rotation invariant (in particular single-valued at the poles, unlike white noise per grid cell).

This module restates that construction on the GPU: geodesy-normalised associated Legendre functions
by the standard stable recurrence (float64, host, once per grid), then per draw Gaussian coefficients
(torch's device generator) and gc_sh_synthesis, the library's hand-written Legendre + longitude
synthesis, which writes the sampler's [member * node, channel] state layout directly.  There is no
CPU path; the float64 host restatement used by the tests is oracle/spherical_noise_oracle.py.
dinosaur is not installable here, so agreement with its exact coefficient ordering / random stream is
unpinned; what is tested is the synthesis against the float64 restatement on the same
coefficients and the statistical contract (zero mean, unit variance at every latitude, isotropy).
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import ops


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


def amplitude_table(grid_lat, n_lon: int) -> np.ndarray:
    """table[m, l, lat] = Pbar[m, l, lat] * sqrt(power_l / (2 l + 1)) with the flat spectrum power_l = 1 / (n_lon / 2)
    of the reference (gencast/samplers_utils.py:316-322, :336-344), float64."""
    L = max(1, n_lon // 2)
    table = legendre_table(L, np.sin(np.deg2rad(np.asarray(grid_lat, np.float64))))
    amp = np.sqrt((1.0 / L) / (2.0 * np.arange(L) + 1.0))
    return table * amp[None, :, None]


class SphericalNoise:
    """Unit-variance isotropic white noise fields on an equiangular lat/lon grid (poles included), on the GPU."""

    def __init__(self, grid_lat, grid_lon, device=None):
        self.device = torch.device(device if device is not None else "cuda:0")
        if self.device.type != "cuda":
            raise RuntimeError("SphericalNoise runs on a CUDA device (gc_sh_synthesis); the float64 host restatement "
                               "for tests is oracle/spherical_noise_oracle.py")
        self.n_lat, self.n_lon = len(grid_lat), len(grid_lon)
        self.L = max(1, self.n_lon // 2)                     # gencast/samplers_utils.py:336
        self.table = torch.from_numpy(amplitude_table(grid_lat, self.n_lon).astype(np.float32)).to(self.device)   # [m, l, lat]
        self._spec = None

    def draw_coefficients(self, n_fields: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """[2 (cos | sin), m, field, l] standard normal coefficients (entries with l < m are not used)."""
        return torch.randn(2, self.L, n_fields, self.L, generator=generator, device=self.device)

    def synthesize(self, coef: torch.Tensor, channels: int, members: int = 1) -> torch.Tensor:
        """Coefficients [2, L, members * channels, L] -> [members * n_lat * n_lon, channels] fp32 (the sampler's state
        layout: node index = lat * n_lon + lon, member-major blocks)."""
        need = members * self.n_lat * channels * self.L * 2
        if self._spec is None or self._spec.numel() < need:
            self._spec = torch.empty(need, dtype=torch.float32, device=self.device)
        out = torch.empty(members * self.n_lat * self.n_lon, channels, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            return ops.sh_synthesis(coef, self.table, self._spec, out, members, channels, self.n_lat, self.n_lon)

    def sample_nodes(self, channels: int, members: int = 1, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """[members * n_lat * n_lon, channels] unit-variance noise."""
        return self.synthesize(self.draw_coefficients(members * channels, generator), channels, members)

    def sample(self, n_fields: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """[n_fields, n_lat, n_lon] fp32 on the device."""
        nodes = self.sample_nodes(n_fields, 1, generator)
        return nodes.reshape(self.n_lat, self.n_lon, n_fields).permute(2, 0, 1).contiguous()
