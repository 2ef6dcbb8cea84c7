"""Minimal stand-in for flax.nnx.Rngs at the API edge (the reference passes
`nnx.Rngs` into GenCast / Sampler: gencast/gencast.py:154, dpm_solver_plus_plus_2s.py:52-56).

Only the behaviour the hot path uses is kept: named streams that hand out a fresh
key each time they are called.  Keys are 63-bit integers that seed the device
noise generator; the draws are not bit-compatible with JAX's threefry streams.
"""
from __future__ import annotations

import numpy as np


class RngStream:
    def __init__(self, seed: int):
        self._seq = np.random.SeedSequence(int(seed))

    def __call__(self) -> int:
        child = self._seq.spawn(1)[0]
        return int(child.generate_state(1, dtype=np.uint64)[0] >> np.uint64(1))


class Rngs:
    """Rngs(0) or Rngs(noise=1, params=2); `rngs.noise()` returns a new key per call."""

    def __init__(self, default: int = 0, **streams: int):
        self._default = int(default)
        self._streams = {k: RngStream(v) for k, v in streams.items()}

    def __getattr__(self, name: str) -> RngStream:
        if name.startswith("_"):
            raise AttributeError(name)
        if name not in self._streams:
            # derive a distinct stream per name from the default seed
            h = int(np.random.SeedSequence([self._default, *name.encode()]).generate_state(1, dtype=np.uint64)[0] >> np.uint64(1))
            self._streams[name] = RngStream(h)
        return self._streams[name]


def split(rng):
    """(rng, this_rng) like jax.random.split for integer keys (common/rollout.py:307-314)."""
    ss = np.random.SeedSequence(int(rng))
    a, b = ss.spawn(2)
    f = lambda s: int(s.generate_state(1, dtype=np.uint64)[0] >> np.uint64(1))
    return f(a), f(b)
