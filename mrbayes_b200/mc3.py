"""ctypes view of the MC^3 shard coordinator (include/mb200_mc3.h, mrbayes_b200/lib/libmb200_mc3.so).

Test and benchmark tooling only: the coordinator itself is C (chain -> process map, the
{lnL, lnPrior, chainId} all-gather over NCCL, the swap rule, the end-of-run reduce).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
MC3_LIB = ROOT / "mrbayes_b200" / "lib" / "libmb200_mc3.so"
ID_BYTES = 128
NCCL, LOOPBACK = 0, 1
ROW = 3                      # doubles per chain in the gathered table: lnL, lnPrior, chainId


class Config(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("device", C.c_int), ("num_runs", C.c_int),
                ("chains_per_run", C.c_int), ("num_swaps", C.c_int), ("chain_temp", C.c_double),
                ("swap_seed", C.c_long), ("backend", C.c_int)]


_lib = None


def library():
    global _lib
    if _lib is None:
        if not MC3_LIB.exists():
            raise FileNotFoundError(f"{MC3_LIB} is missing: build it first (python -c 'import __graft_entry__ as g; g.build()')")
        # one NCCL per process: if PyTorch is (or will be) loaded in this process, its bundled libnccl.so.2
        # must be the copy that satisfies our DT_NEEDED entry too (same SONAME, newer than the system's)
        try:
            import importlib.util
            spec = importlib.util.find_spec("nvidia.nccl")
            if spec is not None and spec.submodule_search_locations:
                cand = Path(list(spec.submodule_search_locations)[0]) / "lib" / "libnccl.so.2"
                if cand.exists():
                    C.CDLL(str(cand), mode=C.RTLD_GLOBAL)
        except Exception:
            pass
        L = C.CDLL(str(MC3_LIB))
        V = C.c_void_p
        L.mb200_mc3_unique_id.argtypes = [C.c_char_p]
        L.mb200_mc3_create.argtypes = [C.POINTER(Config), C.c_char_p, C.POINTER(V)]
        L.mb200_mc3_destroy.argtypes = [V]
        for name in ("local_chain_count", "first_local_chain"):
            getattr(L, f"mb200_mc3_{name}").argtypes = [V]
        for name in ("owner", "chain_id"):
            getattr(L, f"mb200_mc3_{name}").argtypes = [V, C.c_int]
        L.mb200_mc3_temperature.argtypes = [V, C.c_int]
        L.mb200_mc3_temperature.restype = C.c_double
        L.mb200_mc3_exchange_begin.argtypes = [V, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.mb200_mc3_exchange_end.argtypes = [V]
        L.mb200_mc3_table.argtypes = [V]
        L.mb200_mc3_table.restype = C.POINTER(C.c_double)
        L.mb200_mc3_attempt_swaps.argtypes = [V, C.POINTER(C.c_int)]
        L.mb200_mc3_next_swaps_cross_ranks.argtypes = [V]
        L.mb200_mc3_reduce_sum.argtypes = [V, C.POINTER(C.c_double), C.c_int, C.c_int]
        L.mb200_mc3_barrier.argtypes = [V]
        L.mb200_mc3_swap_info.argtypes = [V, C.POINTER(C.c_int)]
        L.mb200_mc3_collectives.argtypes = [V]
        L.mb200_mc3_collectives.restype = C.c_longlong
        L.mb200_mc3_decision_hash.argtypes = [V]
        L.mb200_mc3_decision_hash.restype = C.c_ulonglong
        L.mb200_mc3_run_hash.argtypes = [V, C.c_int]
        L.mb200_mc3_run_hash.restype = C.c_ulonglong
        L.mb200_mc3_run_missed.argtypes = [V, C.c_int]
        L.mb200_mc3_run_missed.restype = C.c_longlong
        _lib = L
    return _lib


def unique_id() -> bytes:
    buf = C.create_string_buffer(ID_BYTES)
    rc = library().mb200_mc3_unique_id(buf)
    if rc != 0:
        raise RuntimeError(f"mb200_mc3_unique_id failed with code {rc}")
    return buf.raw


class Coordinator:
    def __init__(self, rank=0, world=1, device=0, num_runs=2, chains_per_run=4, num_swaps=1, chain_temp=0.1,
                 swap_seed=12345, backend=NCCL, nccl_id: bytes | None = None):
        self.lib = library()
        self.cfg = Config(rank, world, device, num_runs, chains_per_run, num_swaps, chain_temp, swap_seed, backend)
        self.handle = C.c_void_p()
        rc = self.lib.mb200_mc3_create(C.byref(self.cfg), nccl_id, C.byref(self.handle))
        if rc != 0:
            raise RuntimeError(f"mb200_mc3_create failed with code {rc}")
        self.n_global = num_runs * chains_per_run
        self.n_local = self.lib.mb200_mc3_local_chain_count(self.handle)
        self.first = self.lib.mb200_mc3_first_local_chain(self.handle)

    def close(self):
        if self.handle:
            self.lib.mb200_mc3_destroy(self.handle)
            self.handle = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, name, rc):
        if rc != 0:
            raise RuntimeError(f"mb200_mc3_{name} failed with code {rc}")

    def owner(self, g):
        return self.lib.mb200_mc3_owner(self.handle, g)

    def chain_id(self, g):
        return self.lib.mb200_mc3_chain_id(self.handle, g)

    def temperature(self, g):
        return self.lib.mb200_mc3_temperature(self.handle, g)

    def exchange_begin(self, lnl, lnprior):
        a = np.ascontiguousarray(lnl, np.float64)
        b = np.ascontiguousarray(lnprior, np.float64)
        assert a.shape == b.shape == (self.n_local,)
        self._check("exchange_begin", self.lib.mb200_mc3_exchange_begin(
            self.handle, a.ctypes.data_as(C.POINTER(C.c_double)), b.ctypes.data_as(C.POINTER(C.c_double))))

    def exchange_end(self):
        self._check("exchange_end", self.lib.mb200_mc3_exchange_end(self.handle))

    def table(self) -> np.ndarray:
        """The gathered table as a writable view [n_global][3]."""
        p = self.lib.mb200_mc3_table(self.handle)
        return np.ctypeslib.as_array(p, shape=(self.n_global, ROW))

    def attempt_swaps(self) -> int:
        n = C.c_int(0)
        self._check("attempt_swaps", self.lib.mb200_mc3_attempt_swaps(self.handle, C.byref(n)))
        return n.value

    def next_swaps_cross_ranks(self) -> bool:
        return bool(self.lib.mb200_mc3_next_swaps_cross_ranks(self.handle))

    def reduce_sum(self, values, root=0):
        v = np.ascontiguousarray(values, np.float64).copy()
        self._check("reduce_sum", self.lib.mb200_mc3_reduce_sum(self.handle, v.ctypes.data_as(C.POINTER(C.c_double)), v.size, root))
        return v

    def barrier(self):
        self._check("barrier", self.lib.mb200_mc3_barrier(self.handle))

    def swap_info(self) -> np.ndarray:
        n = self.cfg.chains_per_run
        out = np.zeros((self.cfg.num_runs, n, n), np.int32)
        self._check("swap_info", self.lib.mb200_mc3_swap_info(self.handle, out.ctypes.data_as(C.POINTER(C.c_int))))
        return out

    def collectives(self) -> int:
        return self.lib.mb200_mc3_collectives(self.handle)

    def decision_hash(self) -> int:
        return self.lib.mb200_mc3_decision_hash(self.handle)

    def run_hash(self, run) -> int:
        return self.lib.mb200_mc3_run_hash(self.handle, run)

    def run_missed(self, run) -> int:
        return self.lib.mb200_mc3_run_missed(self.handle, run)
