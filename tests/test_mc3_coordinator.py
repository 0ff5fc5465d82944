"""MC^3 shard coordinator (include/mb200_mc3.h) on CPU.

* the swap rule against a direct restatement of the reference's AttemptSwap / GetSwappers /
  Temperature / RandomNumber (src/mcmc.c:591-760, 5213-5246, 18963; src/utils.c:13802);
* the chain -> process map of SetLocalChainsAndDataSplits (src/mcmc.c:18331);
* world-size independence: the same seed and the same chains give the same swap decisions whether the
  chains live on 1, 2, 4 or 8 processes -- first with several coordinators in one process (LOOPBACK
  backend, rows copied by hand), then with two real processes over gloo (the GPU run moves the very
  same rows with ncclAllGather).
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from mrbayes_b200 import mc3  # noqa: E402

pytestmark = pytest.mark.skipif(not mc3.MC3_LIB.exists(), reason="libmb200_mc3.so not built")


# ---- the reference's rules, restated independently (pure Python) ---------------------------------
def park_miller(seed):
    hi, lo = divmod(seed, 127773)
    test = 16807 * lo - 2836 * hi
    seed = test if test > 0 else test + 2147483647
    return seed, seed / 2147483647.0


def reference_swaps(lnl_by_gen, lnpr_by_gen, runs, chains, num_swaps, chain_temp, seed):
    """-> list of (a, b, accepted) and final chainId, applying the serial reference's rules per generation."""
    chain_id = list(range(runs * chains))
    out = []
    for lnl, lnpr in zip(lnl_by_gen, lnpr_by_gen):
        for run in range(runs):
            for _ in range(num_swaps):
                seed, u1 = park_miller(seed)
                seed, u2 = park_miller(seed)
                a = int(u1 * chains)
                b = int(u2 * (chains - 1))
                if b == a:
                    b = chains - 1
                a += run * chains
                b += run * chains
                seed, u = park_miller(seed)
                tA = 1.0 / (1.0 + chain_temp * (chain_id[a] % chains))
                tB = 1.0 / (1.0 + chain_temp * (chain_id[b] % chains))
                lnR = (tB * (lnl[a] + lnpr[a]) + tA * (lnl[b] + lnpr[b])) - (tA * (lnl[a] + lnpr[a]) + tB * (lnl[b] + lnpr[b]))
                r = 0.0 if lnR < -100.0 else 1.0 if lnR > 0.0 else float(np.exp(lnR))
                ok = u < r
                if ok:
                    chain_id[a], chain_id[b] = chain_id[b], chain_id[a]
                out.append((a, b, ok))
    return out, chain_id


def synthetic_trajectory(n_global, gens, seed):
    rng = np.random.default_rng(seed)
    lnl = -5000.0 + np.cumsum(rng.normal(0, 3.0, size=(gens, n_global)), axis=0)
    lnpr = rng.normal(-20, 2.0, size=(gens, n_global))
    return lnl, lnpr


def run_world(world, runs, chains, num_swaps, temp, seed, lnl, lnpr):
    """`world` coordinators in one process; the 'transport' is a plain row copy."""
    cs = [mc3.Coordinator(rank=r, world=world, num_runs=runs, chains_per_run=chains, num_swaps=num_swaps,
                          chain_temp=temp, swap_seed=seed, backend=mc3.LOOPBACK) for r in range(world)]
    try:
        accepted = 0
        for g in range(lnl.shape[0]):
            for c in cs:
                c.exchange_begin(lnl[g, c.first:c.first + c.n_local], lnpr[g, c.first:c.first + c.n_local])
            rows = np.concatenate([c.table()[c.first:c.first + c.n_local].copy() for c in cs])      # the all-gather
            for c in cs:
                c.table()[:] = rows
                c.exchange_end()
            acc = [c.attempt_swaps() for c in cs]
            assert len(set(acc)) == 1
            accepted += acc[0]
        hashes = {c.decision_hash() for c in cs}
        assert len(hashes) == 1, "processes disagree on the swap history"
        ids = [[c.chain_id(g) for g in range(runs * chains)] for c in cs]
        assert all(i == ids[0] for i in ids)
        return hashes.pop(), ids[0], accepted, cs[0].swap_info()
    finally:
        for c in cs:
            c.close()


@pytest.mark.parametrize("runs,chains,num_swaps", [(2, 4, 1), (1, 16, 4), (4, 8, 2)])
def test_swap_rule_matches_reference_restatement(runs, chains, num_swaps):
    n, gens, temp, seed = runs * chains, 300, 0.1, 12345
    lnl, lnpr = synthetic_trajectory(n, gens, 1)
    want, want_ids = reference_swaps(lnl, lnpr, runs, chains, num_swaps, temp, seed)
    with mc3.Coordinator(num_runs=runs, chains_per_run=chains, num_swaps=num_swaps, chain_temp=temp, swap_seed=seed) as c:
        got_acc = 0
        for g in range(gens):
            c.exchange_begin(lnl[g], lnpr[g])
            c.exchange_end()
            got_acc += c.attempt_swaps()
        assert got_acc == sum(ok for _, _, ok in want) > 0
        assert [c.chain_id(g) for g in range(n)] == want_ids
        info = c.swap_info()
        assert info.sum() == len(want) + got_acc                  # lower triangle: attempts, upper: acceptances
        for g in range(n):
            assert c.temperature(g) == pytest.approx(1.0 / (1.0 + temp * (want_ids[g] % chains)))


def test_chain_to_process_map():
    with mc3.Coordinator(rank=2, world=4, num_runs=1, chains_per_run=16, backend=mc3.LOOPBACK) as c:
        assert (c.n_local, c.first) == (4, 8)
        assert [c.owner(g) for g in range(16)] == [g // 4 for g in range(16)]       # contiguous blocks (src/mcmc.c:617-646)
    with pytest.raises(RuntimeError):       # chains must divide evenly over the processes (src/mcmc.c:18338)
        mc3.Coordinator(rank=0, world=3, num_runs=2, chains_per_run=4, backend=mc3.LOOPBACK)
    with pytest.raises(RuntimeError):       # at least as many chains as processes (src/mcmc.c:18356)
        mc3.Coordinator(rank=0, world=16, num_runs=2, chains_per_run=4, backend=mc3.LOOPBACK)


@pytest.mark.parametrize("runs,chains,num_swaps", [(1, 16, 2), (2, 8, 1), (1, 32, 4)])
def test_swap_decisions_do_not_depend_on_world_size(runs, chains, num_swaps):
    n, gens, temp, seed = runs * chains, 200, 0.05, 777
    lnl, lnpr = synthetic_trajectory(n, gens, 2)
    base = run_world(1, runs, chains, num_swaps, temp, seed, lnl, lnpr)
    assert base[2] > 0
    for world in (2, 4, 8):
        got = run_world(world, runs, chains, num_swaps, temp, seed, lnl, lnpr)
        assert got[0] == base[0] and got[1] == base[1] and got[2] == base[2]
        assert np.array_equal(got[3], base[3])


def test_no_collective_when_swappers_are_co_resident():
    """Whole runs per process: every swap pair lives on one process, nothing needs to travel
    (the reference's `procIdForA == procIdForB` branch, src/mcmc.c:668)."""
    with mc3.Coordinator(rank=0, world=2, num_runs=4, chains_per_run=4, backend=mc3.LOOPBACK) as c:
        for _ in range(50):
            assert not c.next_swaps_cross_ranks()
            c.exchange_begin(np.full(8, -100.0), np.zeros(8)); c.exchange_end(); c.attempt_swaps()
    with mc3.Coordinator(rank=0, world=4, num_runs=1, chains_per_run=16, backend=mc3.LOOPBACK) as c:
        crossing = 0
        for _ in range(50):
            crossing += c.next_swaps_cross_ranks()
            c.exchange_begin(np.full(4, -100.0), np.zeros(4)); c.exchange_end(); c.attempt_swaps()
        assert crossing > 25          # P(cross) = 12/15 per swap


# ---- two real processes over gloo ------------------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        runs, chains, num_swaps, temp, seed, gens = 1, 8, 2, 0.1, 4242, 150
        lnl, lnpr = synthetic_trajectory(runs * chains, gens, 3)
        c = mc3.Coordinator(rank=rank, world=world, num_runs=runs, chains_per_run=chains, num_swaps=num_swaps,
                            chain_temp=temp, swap_seed=seed, backend=mc3.LOOPBACK)
        for g in range(gens):
            c.exchange_begin(lnl[g, c.first:c.first + c.n_local], lnpr[g, c.first:c.first + c.n_local])
            mine = torch.from_numpy(c.table()[c.first:c.first + c.n_local].copy())
            rows = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(rows, mine)                                   # what ncclAllGather does on the GPUs
            c.table()[:] = torch.cat(rows).numpy()
            c.exchange_end()
            c.attempt_swaps()
        want, want_ids = reference_swaps(lnl, lnpr, runs, chains, num_swaps, temp, seed)
        assert [c.chain_id(g) for g in range(runs * chains)] == want_ids
        hashes = [None] * world
        dist.all_gather_object(hashes, c.decision_hash())
        assert len(set(hashes)) == 1
        # end of run: one double per run summed over the processes (MPI_Reduce, src/mcmc.c:17246)
        t = torch.tensor([float(lnl[-1, c.first:c.first + c.n_local].sum())], dtype=torch.float64)
        dist.reduce(t, dst=0)
        if rank == 0:
            assert t.item() == pytest.approx(float(lnl[-1].sum()), rel=1e-12)
        c.close()
        Path(out_dir, f"ok{rank}").write_text(str(hashes[0]))
    finally:
        dist.destroy_process_group()


def test_swap_exchange_over_two_gloo_ranks(tmp_path):
    pytest.importorskip("torch")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    h = {(tmp_path / f"ok{r}").read_text() for r in range(world)}
    assert len(h) == 1
    # and the single-process answer for the same seed
    lnl, lnpr = synthetic_trajectory(8, 150, 3)
    with mc3.Coordinator(num_runs=1, chains_per_run=8, num_swaps=2, chain_temp=0.1, swap_seed=4242) as c:
        for g in range(150):
            c.exchange_begin(lnl[g], lnpr[g]); c.exchange_end(); c.attempt_swaps()
        assert str(c.decision_hash()) == h.pop()
