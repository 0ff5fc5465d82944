"""Tuning probe (not a test): primates step time and nuc200k evaluation time for the current
MB200_NT_SMALL / MB200_NT_STREAM environment."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench
from mrbayes_b200 import abi
lib = abi.engine_library()
pr = bench.primates_problem(8, 1)
inst = pr.create(lib, max_evaluations=8)
steps = bench.make_cycle(pr, inst, 128, 3)
batches = [inst.pack(s) for s in steps]
stream = torch.cuda.ExternalStream(inst.stream())
for i in range(128): inst.replay(batches[i])
inst.synchronize()
ms = bench.time_resident(torch, inst, stream, batches, [i % 128 for i in range(2048)], None)
upd = sum(bench.updates_of(s, pr.C, pr.K) for s in steps) * 16
print(f"primates warm: {ms/2048*1e3:.2f} us/step  {upd/(ms*1e-3):.3e} upd/s")
inst.close()
if len(sys.argv) > 1:
    big = bench.synthetic_problem("nuc200k", 1, 2026)
    with big.create(lib) as bi:
        bi.evaluate(big.full_evaluation(0))
        b = bi.pack([big.full_evaluation(0)])
        st = torch.cuda.ExternalStream(bi.stream())
        for _ in range(3): bi.replay(b)
        bi.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(10): bi.replay(b)
        e.record(st); bi.synchronize()
        t = a.elapsed_time(e) / 10
        u = big.n_int * big.C * big.K
        print(f"nuc200k: {t:.4f} ms/eval  {u/(t*1e-3):.3e} upd/s  frac {u*51/(t*1e-3)/6569.6e9:.3f}")
