"""Tiny driver for ncu: full-tree evaluations of the large synthetic 4-state workload
(200 000 patterns x 4 rates, 32 taxa): the bandwidth-bound kernel."""
import sys
sys.path.insert(0, '/root/repo')
import bench
from mrbayes_b200 import abi
lib = abi.engine_library()
name = sys.argv[1] if len(sys.argv) > 1 else "nuc200k"
pr = bench.synthetic_problem(name, 1, 2026)
inst = pr.create(lib)
for it in range(4):
    inst.evaluate(pr.full_evaluation(0))
inst.close()
