"""CPU, world_size 2 over gloo: the N>1 exchange (16-byte argmax record + score gather) of bode_b200.dist."""
import os
import socket

import numpy as np
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from bode_b200 import dist


class _FakeEngine:
    """Stands in for EkldEngine on CPU: 'scores' a shard with a fixed function so the collective logic is testable."""
    device = torch.device("cpu")

    def __init__(self, table):
        self.table = table

    def score(self, Xd, idx_offset=0, form=0, want_var=True):
        idx = (Xd[:, 0] * 1e6).round().astype(np.int64)
        sc = torch.from_numpy(self.table[idx])
        j = int(np.argmax(self.table[idx]))
        return dict(score=sc, var=sc * 0 + 1.0, best=sc[j:j + 1].clone(), best_idx=torch.tensor([idx_offset + j]))


def _worker(rank, world, port, table, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = table.shape[0]
        Xd = (np.arange(n)[:, None] / 1e6)
        res = dist.sharded_score(_FakeEngine(table), Xd, group=tdist.group.WORLD)  # shards only with an explicit group
        unsharded = dist.sharded_score(_FakeEngine(table), Xd)
        assert unsharded["idx_best"] == res["idx_best"] and len(unsharded["score"]) == n
        q.put((rank, res["idx_best"], res["best"], res["score"].tolist()))
    finally:
        tdist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _run(table, world=2):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, table, q)) for r in range(world)]
    [p.start() for p in procs]
    out = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    return sorted(out)


def test_two_rank_argmax_matches_numpy():
    rng = np.random.default_rng(5)
    table = rng.normal(size=101)
    table[77] = table.max() + 1.0
    out = _run(table)
    for rank, idx, best, score in out:
        assert idx == 77 and best == table[77]
        np.testing.assert_array_equal(np.array(score), table)  # gathered full score vector, identical on all ranks


def test_two_rank_tie_and_nan_semantics():
    table = np.zeros(10)
    table[[3, 8]] = 5.0  # tie across the two shards -> lowest global index
    assert all(o[1] == 3 for o in _run(table))
    table[9] = np.nan     # a NaN on rank 1 wins, like np.argmax
    assert all(o[1] == 9 for o in _run(table))


# ---- a sharded KLSampler.optimize keeps its ranks consistent (every rank used to draw its own design) --------
def _optimize_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bode_b200
        from tests import cases
        from tests.fake_engine import OracleEngine
        np.random.seed(100 + rank)  # deliberately different host RNG streams per rank
        X0 = np.array([[0.1, 0.2], [0.8, 0.3], [0.4, 0.9], [0.6, 0.6]])
        f = lambda x: float(np.sin(3 * x[0]) + x[1] ** 2 + 0.05 * np.random.randn())  # noisy objective: rank 0's draw must win
        Y0 = np.array([[np.sin(3 * x[0]) + x[1] ** 2] for x in X0])

        def hyper(X, Y, it):  # stands in for the host MCMC: uses the global RNG, so it differs per rank unless broadcast
            return np.column_stack([np.random.gamma(2.0, 1.0, 5) + 0.2, np.random.uniform(0.3, 0.9, 5), np.random.uniform(0.3, 0.9, 5)])

        kls = bode_b200.KLSampler(X0, Y0, False, f, False, [(0, 1)] * 2, mcmc_model=True, max_it=3, kld_tol=0.0, mcmc_chains=2,
                                  mcmc_model_avg=4, mcmc_steps=40, mcmc_burn=10, mcmc_thin=5, engine=OracleEngine(),
                                  hyper_samples=hyper, ekld_form=0, process_group=tdist.group.WORLD)
        X, Y, X_u, kld_all, X_design, mu, sg, comp = kls.optimize(num_designs=60, comp=True)
        q.put((rank, X.tolist(), Y.tolist(), X_u.tolist(), X_design.tolist(), kls.chosen_idx, kls.model[0].tolist(), kld_all.tolist()))
    finally:
        tdist.destroy_process_group()


def test_two_rank_optimize_ranks_stay_consistent():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_optimize_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted(q.get(timeout=180) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    a, b = out
    assert a[1:] == b[1:]            # observations, comparator data, last design, chosen indices, hyper-samples, scores
    assert len(a[5]) == 3 and len(a[1]) == 4 + 3
