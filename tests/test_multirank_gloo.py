"""N > 1 path of bench.py on CPU: two processes over gloo (the GPU run moves the same rows with NCCL).

ONE run of primates (nruns=1, nchains=8) whose heated chains are dealt out over the processes in
contiguous blocks (the reference's chain -> process map, src/mcmc.c:18331): every generation each
process evaluates its own chains (CPU oracle here: test infrastructure), then {lnL, lnPrior, chainId}
of all chains is all-gathered and every process attempts the same swaps.  The swap decisions and the
final heats must be what ONE process holding all eight chains gets from the same seed -- a chain's
trajectory and the swap generator do not depend on the number of processes."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

GENS = 24


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _single_process_answer():
    import bench
    from mrbayes_b200 import abi, mc3
    w = bench.WORKLOADS["primates-sharded"]
    job = bench.Job("primates-sharded", 0, 1, abi.oracle_library(), 0, 16)
    with mc3.Coordinator(num_runs=job.runs, chains_per_run=job.chains, num_swaps=w["swaps"], chain_temp=0.1, swap_seed=12345) as mc:
        cur = bench.python_mc3_loop(job, mc, GENS)
        out = (mc.decision_hash(), [mc.chain_id(g) for g in range(job.runs * job.chains)], cur.copy(), int(mc.swap_info().sum()))
    job.close()
    return out


def _worker(rank: int, world: int, port: int, out_dir: str):
    import torch
    import torch.distributed as dist
    import bench
    from mrbayes_b200 import abi, mc3

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = bench.WORKLOADS["primates-sharded"]
        job = bench.Job("primates-sharded", rank, world, abi.oracle_library(), 0, 16)
        assert job.n_local == 4 and job.globals == list(range(4 * rank, 4 * rank + 4))
        mc = mc3.Coordinator(rank=rank, world=world, num_runs=job.runs, chains_per_run=job.chains, num_swaps=w["swaps"],
                             chain_temp=0.1, swap_seed=12345, backend=mc3.LOOPBACK)

        def gather(rows):
            mine = torch.from_numpy(rows)
            parts = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            return torch.cat(parts).numpy()

        cur = bench.python_mc3_loop(job, mc, GENS, gather)
        # end of run: one double per run summed over the processes (MPI_Reduce, src/mcmc.c:17246)
        t = torch.tensor([float(cur.sum())], dtype=torch.float64)
        dist.all_reduce(t)
        ids = [mc.chain_id(g) for g in range(job.runs * job.chains)]
        np.save(Path(out_dir, f"r{rank}.npy"), np.array([mc.decision_hash() & 0xffffffff, mc.decision_hash() >> 32, *ids, t.item(), *cur]))
        mc.close(); job.close()
    finally:
        dist.destroy_process_group()


def test_one_run_sharded_over_two_ranks_matches_single_process(tmp_path):
    pytest.importorskip("torch")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"r{k}.npy") for k in range(world)]
    h1, ids1, cur1, nswap = _single_process_answer()
    assert nswap > GENS                                   # swaps were attempted (2 per generation) and some accepted
    for k in range(world):
        assert (int(r[k][0]) | (int(r[k][1]) << 32)) == h1, "swap decisions differ from the single-process run"
        assert [int(x) for x in r[k][2:10]] == ids1
        assert r[k][10] == pytest.approx(float(cur1.sum()), rel=1e-12)
        assert np.array_equal(r[k][11:], cur1[4 * k:4 * k + 4])       # a chain's trajectory does not depend on its owner
