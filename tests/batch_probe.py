"""Dev probe: ONE instance holding many chains (nruns x nchains of the same alignment), all of a generation
in ONE launch -- against the same number of chains spread over independent 8-chain instances."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from mrbayes_b200 import abi
lib = abi.engine_library()
for nch in (8, 64, 128, 256):
    pr = bench.primates_problem(nch, seed=11)
    inst = pr.create(lib, max_evaluations=nch)
    steps = bench.make_cycle(pr, inst, 16, 5)
    batches = [inst.pack(s) for s in steps]
    stream = torch.cuda.ExternalStream(inst.stream())
    upd = sum(bench.updates_of(s, pr.C, pr.K) for s in steps)
    for rep in range(3):
        for b in batches: inst.replay(b)
    inst.synchronize()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for rep in range(20):
        for b in batches: inst.replay(b)
    e.record(stream); inst.synchronize()
    ms = a.elapsed_time(e) / (20 * len(batches))
    print(f"{nch:4d} chains in one launch: {ms*1e3:7.1f} us/step, {upd/len(batches)/(ms*1e-3):.3e} upd/s (warm L2)")
    inst.close()
