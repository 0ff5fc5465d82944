import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
from mrbayes_b200 import abi
import bench
lib = abi.Library('/root/repo/mrbayes_b200/lib/libmb200_dbg.so','mb200_')
lib.fn('debug_read_stamps').argtypes=[C.c_int, C.POINTER(C.c_ulonglong), C.c_int]
for name in ("codon20k","aa50k"):
    pr = bench.synthetic_problem(name, 1, 2026)
    inst = pr.create(lib)
    inst.evaluate(pr.full_evaluation(0))
    sp = pr.full_evaluation(0)
    inst.evaluate(sp)
    buf=(C.c_ulonglong*64)()
    lib.fn('debug_read_stamps')(inst.handle, buf, 1)
    r=np.array(buf[:],dtype=np.uint64).astype(np.int64)
    op=sp.ops[5]
    print(name, "op5 children", op['child1'], op['child2'], "tips<", pr.n_tips)
    for ch in range(2):
        b=8+ch*8
        print(f"  child{ch}: A(load+split)={r[b+1]-r[b]} waitB={r[b+2]-r[b+1]} sync+issue={r[b+3]-r[b+2]} mma={r[b+4]-r[b+3]} epi1={r[b+5]-r[b+4]} sync={r[b+6]-r[b+5]}")
    print(f"  epi2: scale={r[44]-r[40]} stage+sync={r[45]-r[44]} copy={r[46]-r[45]} sync={r[41]-r[46]}  op0->op1 total={r[43]-r[42]}")
    inst.close()
