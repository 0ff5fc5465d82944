"""N > 1 path of bench.py on CPU: two processes over gloo (the GPU run uses the same code over NCCL).

Every rank drives its own independent analyses (weak scaling, no data-path collective); what crosses
ranks is (1) the bench contract's reduction -- times MAX over ranks, work SUM over ranks -- and
(2) one all-reduce of per-run lnL sums (the marginal-likelihood reduce of the reference's MPI
build, src/mcmc.c:17246).  The likelihoods here come from the CPU oracle (test infrastructure)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    import torch
    import torch.distributed as dist
    import bench
    from mrbayes_b200 import abi

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # disjoint replicas per rank
        seeds = bench.replica_seeds(rank, 2)
        gathered = [None] * world
        dist.all_gather_object(gathered, seeds)
        flat = [s for per_rank in gathered for s in per_rank]
        assert len(set(flat)) == len(flat)

        # each rank evaluates one generation of its own analyses (oracle), 8 chains each
        lib = abi.oracle_library()
        lnl_sum, updates = 0.0, 0
        for prob_seed, cycle_seed in seeds:
            pr = bench.primates_problem(8, seed=prob_seed)
            with pr.create(lib, max_evaluations=8) as inst:
                steps = bench.make_cycle(pr, inst, 4, seed=cycle_seed)
                for sp in steps:
                    lnl, st = inst.evaluate(sp)
                    assert not st.any() and np.isfinite(lnl).all()
                    updates += bench.updates_of(sp, pr.C, pr.K)
                lnl_sum += float(lnl.sum())

        # (2) the one data-independent collective: sum of per-run lnL over ranks
        t = torch.tensor([lnl_sum], dtype=torch.float64)
        dist.all_reduce(t)
        all_sums = [None] * world
        dist.all_gather_object(all_sums, lnl_sum)
        assert abs(t.item() - sum(all_sums)) <= 1e-9 * abs(sum(all_sums))

        # (1) the bench contract: MAX of times, SUM of work
        my_ms = 10.0 + 5.0 * rank
        ms_value, ms_warm, ms_e2e, all_updates, all_launches = bench.reduce_over_ranks(
            torch, dist, "cpu", my_ms, my_ms + 1.0, my_ms + 2.0, updates, 3 + rank)
        all_upd = [None] * world
        dist.all_gather_object(all_upd, updates)
        assert ms_value == 10.0 + 5.0 * (world - 1) and ms_warm == ms_value + 1.0 and ms_e2e == ms_value + 2.0
        assert all_updates == float(sum(all_upd)) and all_launches == sum(3 + r for r in range(world))
        Path(out_dir, f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_two_ranks_over_gloo(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
