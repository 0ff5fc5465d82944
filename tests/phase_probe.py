import sys, ctypes as C, numpy as np
sys.path.insert(0,'/root/repo')
from mrbayes_b200 import abi
import bench
lib = abi.Library('/root/repo/mrbayes_b200/lib/libmb200_%s.so' % (sys.argv[1] if len(sys.argv)>1 else 'dbg'),'mb200_')
lib.fn('debug_read_stamps').argtypes=[C.c_int, C.POINTER(C.c_ulonglong), C.c_int]
pr = bench.primates_problem(8, 1)
inst = pr.create(lib, max_evaluations=8)
steps = bench.make_cycle(pr, inst, 16, 3)
import torch
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
for rep in range(2):
  for si in range(16):
    specs = steps[si]
    if rep==1: flush.zero_(); torch.cuda.synchronize()
    inst.evaluate(specs)
    if si not in (15, 3): continue
    buf=(C.c_ulonglong*(8*64))()
    lib.fn('debug_read_stamps')(inst.handle, buf, 8)
    a=np.array(buf[:],dtype=np.uint64).reshape(8,64).astype(np.int64)
    t0=a[:,0].min()
    print("rep",rep,"step",si)
    for e in (1, 7):
        mins=[]
        for o in specs[e].ops:
            sc = inst.get_scalers(int(o['scale_write'])) if o['scale_write'] >= 0 else None
            mins.append((round(float(sc.min()),1), round(float(sc.max()),1)) if sc is not None else None)
        print("  eval", e, "node scaler (min,max) of ln m:", mins)
    for e in range(8):
        nop=len(specs[e].ops); nm=len(specs[e].mats)
        ntips = pr.n_tips
        kinds=[]; prev=-9
        for o in specs[e].ops:
            k=''
            for ch in (o['child1'],o['child2'],o['child3']):
                if ch < 0: continue
                k += 'T' if ch < ntips else ('F' if ch == prev else 'L')
            kinds.append(k); prev=o['dest']
        r=a[e]; ops=[int(r[8+o]-r[3]) if o==0 else int(r[8+o]-r[8+o-1]) for o in range(nop)]
        print(f" eval{e} nOp={nop:2d} nMat={nm:2d} start+{r[0]-t0:6d} hdr={r[1]-r[0]:5d} stage={r[2]-r[1]:5d} exp={r[50]-r[2]:5d} rows={r[51]-r[50]:5d} tab={r[3]-r[51]:5d} ops={ops} kinds={kinds} scal={r[4]-r[8+nop-1] if nop else 0:5d} tail={r[5]-r[4]:5d} fin={r[6]-r[5]:5d} total={r[6]-r[0]:6d} | node1: wait={r[40]-r[8]} mv={r[41]-r[40]} scale={r[42]-r[41]} store={r[43]-r[42]} rest={r[9]-r[43]}")
