"""Multi-rank logic on CPU (gloo, world_size 2): ensemble statistics equal a single-process numpy
formula, members are assigned like the reference's pmap fan-out."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, members, truth, weights, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gencast_flax_nnx_b200 import parallel
    mine = parallel.member_assignment(members.shape[0], world, rank)
    stats = parallel.EnsembleStatistics(members.shape[1:], "cpu")
    for m in mine:
        stats.add(torch.from_numpy(members[m]))
    mean, spread, count = stats.finalize()
    crps = parallel.fair_crps(torch.from_numpy(members[mine]), torch.from_numpy(truth), torch.from_numpy(weights))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), mean=mean.numpy(), spread=spread.numpy(), crps=crps.numpy(),
             count=count, mine=np.asarray(mine))
    dist.destroy_process_group()


def test_ensemble_statistics_world_size_2(tmp_path):
    rng = np.random.default_rng(0)
    M, G, C = 6, 37, 5
    members = rng.standard_normal((M, G, C)).astype(np.float32) * 2 + 1
    truth = rng.standard_normal((G, C)).astype(np.float32)
    weights = np.abs(np.cos(np.linspace(-1.5, 1.5, G))).astype(np.float32)
    port = _free_port()
    mp.spawn(_worker, args=(2, port, members, truth, weights, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    assert r0["mine"].tolist() == [0, 1, 2] and r1["mine"].tolist() == [3, 4, 5]
    assert int(r0["count"]) == M
    x = members.astype(np.float64)
    np.testing.assert_allclose(r0["mean"], x.mean(0), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(r0["spread"], x.std(0, ddof=1), rtol=1e-4, atol=1e-5)
    skill = np.abs(x - truth[None]).mean(0)
    pair = sum(np.abs(x[i] - x[j]) for i in range(M) for j in range(i + 1, M)) / (M * (M - 1))
    ref = ((skill - pair) * weights[:, None]).sum(0) / weights.sum()
    np.testing.assert_allclose(r0["crps"], ref, rtol=1e-5, atol=1e-6)
    for k in ("mean", "spread", "crps"):
        np.testing.assert_array_equal(r0[k], r1[k])                 # every rank holds the same statistics


def test_member_assignment_requires_even_split():
    from gencast_flax_nnx_b200 import parallel
    assert parallel.member_assignment(8, 4, 2) == [4, 5]
    with pytest.raises(ValueError):
        parallel.member_assignment(7, 2, 0)
