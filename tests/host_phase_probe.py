"""Dev probe: where the host-call (e2e) time of the primates step goes: pack / launch API / wait.
Needs mrbayes_b200/lib/libmb200_dbg.so (nvcc ... -DMB200_PHASE_TIMING)."""
import sys, time, ctypes as C
sys.path.insert(0, '/root/repo')
from mrbayes_b200 import abi
import bench
lib = abi.Library('/root/repo/mrbayes_b200/lib/libmb200_dbg.so', 'mb200_')
pr = bench.primates_problem(8, 1)
inst = pr.create(lib, max_evaluations=8)
steps = bench.make_cycle(pr, inst, 128, 3)
import numpy as np
arrs = [abi.make_eval_array(s) for s in steps]
out = (C.c_double * 4)()
evaluate = lib.fn('evaluate')
lnl = np.zeros(8); st = np.zeros(8, np.int32)
p_lnl = lnl.ctypes.data_as(C.POINTER(C.c_double)); p_st = st.ctypes.data_as(C.POINTER(C.c_int))
for rep in range(3):
    lib.fn('debug_host_phases')(out)
    t0 = time.perf_counter()
    n = 0
    for it in range(20):
        for a in arrs:
            evaluate(inst.handle, a, 8, p_lnl, p_st)
            n += 1
    dt = time.perf_counter() - t0
    lib.fn('debug_host_phases')(out)
    c = out[3]
    print(f"rep {rep}: wall/call {dt/n*1e6:.2f} us; pack {out[0]/c:.2f} launch {out[1]/c:.2f} wait {out[2]/c:.2f} us (calls {int(c)})")
