"""Host-side logic that needs no GPU: sharding, the pooled-moment algebra, and the world_size-2 all-reduce of the
pooled statistics over the gloo backend (the N>1 path of SURVEY §8e)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from metropolisengine_b200 import parallel
from metropolisengine_b200.engine import adaptation_constants


def test_shard_ranges_partition_the_chain_ids():
    for n, w in ((10, 1), (10, 3), (65536, 8), (1000003, 8), (8, 8)):
        ranges = [parallel.shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        for (a, b), (c, d) in zip(ranges, ranges[1:]):
            assert b == c and b > a
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(3, 0, 8)


def test_adaptation_constants_match_reference_values():
    # SURVEY.md §4 constants row (live reference)
    assert adaptation_constants(1, 0, .3) == (1.0364333894937898, 1, 4.761904761904762)
    assert adaptation_constants(3, 4, .3)[2] == 2.585350473881254
    assert adaptation_constants(1, 64, .3)[2] == 2.261657784893143


@pytest.mark.parametrize("nr,nc", [(2, 0), (3, 4), (0, 2)])
def test_finalize_pooled_recovers_sample_statistics(nr, nc):
    rng = np.random.default_rng(1)
    d = nr + 2 * nc
    a = rng.standard_normal((d, d))
    x = rng.standard_normal((5000, d)) @ a.T + rng.standard_normal(d)
    shift = x[0] * 0.9
    sums = parallel.pooled_moments_reference(x, shift, nr, nc)
    assert sums.shape == (parallel.pool_words(nr, nc),)
    ps = parallel.finalize_pooled(sums, x.shape[0], shift, nr, nc)
    assert np.allclose(ps["mean_real"], x[:, :nr].mean(0))
    if nr:
        assert np.allclose(ps["cov_real"], np.cov(x[:, :nr].T).reshape(nr, nr))
    if nc:
        c = x[:, nr:nr + nc] + 1j * x[:, nr + nc:]
        cm = c - c.mean(0)
        assert np.allclose(ps["mean_complex"], c.mean(0))
        assert np.allclose(ps["cov_complex"], cm.T @ cm.conj() / (len(c) - 1))
        assert np.allclose(ps["observables_mean"][nr:nr + nc], np.abs(c).mean(0))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nr, nc, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)
        d = nr + 2 * nc
        x = rng.standard_normal((4000, d)) * 0.3 + 1.0          # all samples of the job, identical on every rank
        shift = np.full(d, 0.8)
        lo, hi = parallel.shard_range(x.shape[0], rank, world)    # this rank's chains
        local = parallel.pooled_moments_reference(x[lo:hi], shift, nr, nc)
        t = torch.tensor(np.concatenate([local, [hi - lo]]))
        parallel.allreduce_sum_(t)
        ps = parallel.finalize_pooled(t[:-1].numpy(), t[-1].item(), shift, nr, nc)
        full = parallel.finalize_pooled(parallel.pooled_moments_reference(x, shift, nr, nc), x.shape[0], shift, nr, nc)
        ok = all(np.allclose(ps[k], full[k], rtol=1e-12, atol=1e-14) for k in ("mean_real", "cov_real",
                                                                               "cov_complex", "observables_mean"))
        assert parallel.world() == (rank, world)
        out[rank] = bool(ok and ps["count"] == x.shape[0])
    finally:
        dist.destroy_process_group()


def test_pooled_statistics_allreduce_world_size_2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 3, 4, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}
