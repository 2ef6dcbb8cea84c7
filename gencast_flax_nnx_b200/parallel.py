"""Ensemble members across GPUs: one process per GPU, members are independent.

The sampling path needs no collective (the reference fans members out with
`pmap` and never communicates between them, common/rollout.py:109-175).  Collectives are used
only for ensemble statistics, which the reference does not implement (README mention only); they
are defined here:

  mean      = sum_m x_m / M
  spread    = sqrt( (sum_m x_m^2 - M mean^2) / (M - 1) )                      (unbiased)
  fair CRPS = mean_m |x_m - y|  -  sum_{i<j} |x_i - x_j| / (M (M - 1))        (per point)

Sums are accumulated locally (gc_ensemble_accumulate on the GPU) and combined with one
all-reduce of [2, G, C] floats; CRPS gathers the members once and each rank evaluates its slice
of grid points with the sorted-sample identity  sum_{i<j} |x_i - x_j| = sum_k (2k - M + 1) x_(k).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def member_assignment(num_members: int, world_size: int, rank: int):
    """Members of this rank; like the reference, the count must divide evenly (common/rollout.py:110-112)."""
    if num_members % world_size != 0:
        raise ValueError(f"num_members ({num_members}) must be a multiple of the number of devices ({world_size})")
    per = num_members // world_size
    return list(range(rank * per, (rank + 1) * per))


def _world(group) -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class EnsembleStatistics:
    """Running sum / sum of squares of this rank's members, reduced across ranks on demand."""

    def __init__(self, shape, device, group=None):
        self.group = group
        self.acc = torch.zeros((2,) + tuple(shape), dtype=torch.float32, device=device)
        self.local_members = 0

    def add(self, x: torch.Tensor) -> None:
        if x.is_cuda:
            from . import ops
            ops.ensemble_accumulate(x.contiguous(), self.acc[0], self.acc[1])
        else:                                   # host tensors: only the gloo tests of this file's logic
            self.acc[0] += x
            self.acc[1] += x * x
        self.local_members += 1

    def reset(self) -> None:
        self.acc.zero_()
        self.local_members = 0

    def finalize(self, total_members: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor, int]:
        """(mean, spread, total members); every rank gets the same result.  Passing `total_members`
        (when every rank holds the same count) avoids the host round trip for the member count."""
        rank, world = _world(self.group)
        total = self.acc.clone()
        if world > 1:
            dist.all_reduce(total, group=self.group)
        if total_members is None:
            count = torch.tensor([float(self.local_members)], device=self.acc.device)
            if world > 1:
                dist.all_reduce(count, group=self.group)
            m = int(round(float(count.item())))
        else:
            m = int(total_members)
        mean = total[0] / m
        var = (total[1] - m * mean * mean) / max(m - 1, 1)
        return mean, var.clamp_min(0).sqrt(), m


def fair_crps(local_members: torch.Tensor, truth: torch.Tensor, weights: Optional[torch.Tensor] = None,
              group=None) -> torch.Tensor:
    """Fair CRPS per channel, averaged over grid points (optionally weighted, e.g. by cos latitude).

    local_members: [m_local, G, C] this rank's members; truth: [G, C]; weights: [G] or None.
    Returns [C], identical on every rank.
    """
    rank, world = _world(group)
    if world > 1:
        parts = [torch.empty_like(local_members) for _ in range(world)]
        dist.all_gather(parts, local_members.contiguous(), group=group)
        members = torch.cat(parts, dim=0)
    else:
        members = local_members
    M, G, C = members.shape
    lo, hi = (G * rank) // world, (G * (rank + 1)) // world      # this rank's grid slice
    if members.is_cuda and members.dtype == torch.float32 and 2 <= M <= 64:
        # on the GPU: gc_fair_crps (per point, sorted-sample identity in shared memory) + gc_column_sums
        from . import ops
        xs = members[:, lo:hi].reshape(M, -1).contiguous()
        w = None if weights is None else weights[lo:hi].to(torch.float32).contiguous()
        per_point = ops.fair_crps(xs, truth[lo:hi].to(torch.float32).reshape(-1).contiguous(), w, C)
        num = ops.column_sums(per_point.reshape(hi - lo, C)).to(torch.float64)
        den = (torch.full((C,), float(hi - lo), dtype=torch.float64, device=members.device) if w is None
               else w.to(torch.float64).sum().expand(C).clone())
        both = torch.stack([num, den])
        if world > 1:
            dist.all_reduce(both, group=group)
        return (both[0] / both[1]).to(torch.float32)
    x = members[:, lo:hi].to(torch.float64)
    y = truth[lo:hi].to(torch.float64)
    w = torch.ones(hi - lo, dtype=torch.float64, device=x.device) if weights is None else weights[lo:hi].to(torch.float64)
    skill = (x - y[None]).abs().mean(0)
    xs, _ = torch.sort(x, dim=0)
    coef = (2 * torch.arange(M, dtype=torch.float64, device=x.device) - M + 1)[:, None, None]
    pair = (coef * xs).sum(0) / (M * (M - 1)) if M > 1 else torch.zeros_like(skill)
    num = ((skill - pair) * w[:, None]).sum(0)
    den = w.sum().expand(C).clone()
    both = torch.stack([num, den])
    if world > 1:
        dist.all_reduce(both, group=group)
    return (both[0] / both[1]).to(torch.float32)
