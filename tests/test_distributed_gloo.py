"""World-size-2 test of the multi-process path on CPU (gloo): sharding by global id range and
the all-reduce of the statistics.  The product has no CPU compute path, so the local solver is
injected (the CPU oracle); what is under test is the host logic of pdmp_b200.distributed."""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, outdir):
    for p in (ROOT, os.path.join(ROOT, "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import oracle as orc
    import pdmp_b200 as P
    orc.set_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pb = P.models.problem("tcp2d", tspan=(0.0, 10.0))
        n = 1001   # not divisible by the world size: ragged shards
        xc0 = np.tile([0.05, 0.075], (n, 1)) * (1.0 + 0.001 * np.arange(n))[:, None]
        local = lambda problem, alg, **kw: orc.solve(problem, alg.model, algorithm=alg.algorithm, ode=alg.ode, **kw)
        res = P.distributed.solve_sharded(
            pb, P.CHVGPU("tcp2d"), trajectories=n, local_solve=local, n_jumps=200, seed=5, xc0=xc0,
            sample_times=[2.5, 5.0, 10.0], hist=[(0, 0.0, 4.0)], hist_bins=32)
        with open(os.path.join(outdir, f"r{rank}.pkl"), "wb") as f:
            pickle.dump(dict(shard=res.shard, time=res.time, xd=res.xd, offsets=res.offsets, count=res.stat_count,
                             sum=res.stat_sum, sumsq=res.stat_sumsq, hist=res.hist), f)
    finally:
        dist.destroy_process_group()


def test_shard_range():
    from pdmp_b200.distributed import shard_range
    for n, w in [(10, 3), (1001, 2), (7, 8), (0, 4), (10**9, 8)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2
        assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


@pytest.mark.timeout(300)
def test_two_rank_sharding_and_stats_allreduce(orc, P):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), d), nprocs=2, join=True)
        parts = [pickle.load(open(os.path.join(d, f"r{r}.pkl"), "rb")) for r in range(2)]
    pb = P.models.problem("tcp2d", tspan=(0.0, 10.0))
    n = 1001
    xc0 = np.tile([0.05, 0.075], (n, 1)) * (1.0 + 0.001 * np.arange(n))[:, None]
    full = orc.solve(pb, "tcp2d", algorithm="chv", ode="dp5", trajectories=n, n_jumps=200, seed=5, xc0=xc0,
                     sample_times=[2.5, 5.0, 10.0], hist=[(0, 0.0, 4.0)], hist_bins=32)
    assert parts[0]["shard"] == (0, 501) and parts[1]["shard"] == (501, 500)
    # the union of the shards is the single-process ensemble, trajectory for trajectory
    assert np.array_equal(np.concatenate([p["time"] for p in parts]), full.time)
    assert np.array_equal(np.concatenate([p["xd"] for p in parts]), full.xd)
    # every rank holds the GLOBAL statistics
    for p in parts:
        assert np.array_equal(p["count"], full.stat_count) and np.array_equal(p["hist"], full.hist)
        assert np.allclose(p["sum"], full.stat_sum, rtol=1e-12) and np.allclose(p["sumsq"], full.stat_sumsq, rtol=1e-12)
    assert full.stat_count[0, 0] == n
