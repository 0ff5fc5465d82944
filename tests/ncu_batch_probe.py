"""Tiny driver for ncu: 256 chains of the primates workload on one instance, one fused launch per generation."""
import sys
sys.path.insert(0, '/root/repo')
import bench
from mrbayes_b200 import abi
lib = abi.engine_library()
pr = bench.primates_problem(256, seed=11)
inst = pr.create(lib, max_evaluations=256)
steps = bench.make_cycle(pr, inst, 8, 5)
batches = [inst.pack(s) for s in steps]
for rep in range(2):
    for b in batches:
        inst.replay(b)
inst.synchronize()
inst.close()
