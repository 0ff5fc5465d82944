"""GPU probe (not a test): pipeline trace of the warp-specialised tensor-core kernel's first CTA.
Needs mrbayes_b200/lib/libmb200_dbg.so (the engine compiled with -DMB200_PHASE_TIMING).
usage: python tests/tcp_trace_probe.py [codon20k|aa50k ...]"""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mrbayes_b200 import abi  # noqa: E402
import bench  # noqa: E402

import os
lib = abi.Library(str(ROOT / "mrbayes_b200" / "lib" / os.environ.get("MB200_DBG_LIB", "libmb200_dbg.so")), "mb200_")
lib.fn("debug_read_trace").argtypes = [C.c_int, C.POINTER(C.c_ulonglong), C.c_int]
NAMES = ["kind", "tkt", "dec", "pub", "l0w", "l0d", "l1w", "l1d", "mF", "mE", "mI", "eS", "eA", "eT", "eC", "eF"]
for name in sys.argv[1:] or ["codon20k", "aa50k"]:
    S, K, Cn, tips = bench.SYNTH[name]
    from mrbayes_b200 import workloads
    pr = workloads.make_problem(S, K, Cn, tips, 1, seed=2026)
    with pr.create(lib) as inst:
        inst.evaluate(pr.full_evaluation(0))
        buf = (C.c_ulonglong * 4000)()
        lib.fn("debug_read_trace")(inst.handle, buf, 4000)
        batch = inst.pack([pr.full_evaluation(0)])
        inst.replay(batch)
        inst.synchronize()
        lib.fn("debug_read_trace")(inst.handle, buf, 4000)
        r = np.array(buf[:], dtype=np.uint64).astype(np.int64).reshape(250, 16)
        t0 = r[0, 1]
        print(name, "CTA 0: ns since its first ticket;", " ".join(NAMES))
        for i in range(250):
            if r[i, 3] == 0:
                break
            kind, t, oi = r[i, 0] & 0xff, (r[i, 0] >> 8) & 0xffffff, r[i, 0] >> 32
            row = " ".join(f"{(v - t0) if v else -1:7d}" for v in r[i, 1:])
            print(f"  #{i:3d} k{kind} t{t:4d} op{oi:3d} | {row}")
