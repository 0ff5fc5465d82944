"""Tiny driver for ncu: a few chain-batched full evaluations of the primates workload."""
import sys
sys.path.insert(0, '/root/repo')
import bench
from mrbayes_b200 import abi
lib = abi.engine_library()
pr = bench.primates_problem(8, 1)
inst = pr.create(lib, max_evaluations=8)
for it in range(6):
    inst.evaluate([pr.full_evaluation(ch) for ch in range(8)])
inst.close()
