import ctypes as C, numpy as np, subprocess, sys, os
here = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(here, "libumma_probe.so"))
P = C.POINTER(C.c_float)
rng = np.random.default_rng(0)
for (N, KP) in [(64, 64), (32, 24)]:
    A = rng.random((128, KP), dtype=np.float32) + 0.01
    B = rng.random((N, KP), dtype=np.float32) * 0.1
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    for split in (0, 1):
        D = np.zeros((128, N), np.float32)
        rc = lib.umma_probe(A.ctypes.data_as(P), B.ctypes.data_as(P), D.ctypes.data_as(P), N, KP, split)
        err = np.abs(D - ref) / np.abs(ref)
        f32 = (A @ B.T)
        print(f"N={N} KP={KP} split={split} rc={rc} max rel err {err.max():.3e} mean {err.mean():.3e}  (fp32 matmul err {np.abs(f32-ref).max()/np.abs(ref).max():.1e})")
