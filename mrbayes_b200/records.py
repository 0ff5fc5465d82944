"""Reader / replayer for evaluation records ("golden" files).

A record file is what ``oracle/ref_harness.c`` writes in dump mode while the UNMODIFIED
reference runs an MCMC analysis: the exact sequence of engine calls the MrBayes seam would
have issued (instance geometry, tip state sets, pattern weights, eigensystems and one
``EVAL`` record per ``LaunchLogLikeForDivision`` call) together with the log likelihood the
reference's own CPU kernels returned for that call.  Replaying the records through an
``abi.Instance`` (CUDA engine or CPU oracle) reproduces the reference run evaluation by
evaluation, including accept/reject index flips, which are implicit in the buffer indices.

Format: see the header comment of ``oracle/ref_harness.c``.
"""
from __future__ import annotations

import gzip
import struct
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import abi


@dataclass
class Division:
    division: int
    cfg: dict
    tips: dict = field(default_factory=dict)        # tip -> uint64[C]
    weights: dict = field(default_factory=dict)     # row -> float32[C]
    pattern_states: dict | None = None              # variable-state divisions: the 'PSTA' tables


@dataclass
class Event:
    kind: str            # "eigen" | "cijk" | "eval"
    division: int
    eigen: int = -1
    lam: np.ndarray | None = None
    V: np.ndarray | None = None
    Vinv: np.ndarray | None = None
    block: np.ndarray | None = None
    spec: abi.EvalSpec | None = None


CFG_FIELDS = ("tip_count", "partials_count", "state_count", "pattern_count", "category_count",
              "matrix_count", "scaler_count", "eigen_count", "weight_rows", "device",
              "max_evaluations", "flags")


def _open(path):
    path = Path(path)
    if path.suffix == ".gz":
        return gzip.open(path, "rb")
    return open(path, "rb")


def load(path):
    """-> (divisions: dict[int, Division], events: list[Event]) in file order."""
    with _open(path) as f:
        data = f.read()
    if data[:8] != b"MB200GLD":
        raise ValueError(f"{path}: not an mb200 record file")
    (ver,) = struct.unpack_from("<I", data, 8)
    if ver != 1:
        raise ValueError(f"{path}: unsupported version {ver}")
    pos = 12
    divisions: dict[int, Division] = {}
    events: list[Event] = []
    pending_inline = None
    pending_parts = []
    while pos < len(data):
        tag = data[pos:pos + 4]
        (nb,) = struct.unpack_from("<I", data, pos + 4)
        body = memoryview(data)[pos + 8:pos + 8 + nb]
        pos += 8 + nb
        if tag == b"INST":
            vals = struct.unpack_from("<13i", body, 0)
            divisions[vals[0]] = Division(vals[0], dict(zip(CFG_FIELDS, vals[1:])))
        elif tag == b"TIPS":
            d, tip, C = struct.unpack_from("<3i", body, 0)
            divisions[d].tips[tip] = np.frombuffer(body, "<u8", C, 12).copy()
        elif tag == b"WGHT":
            d, row, C = struct.unpack_from("<3i", body, 0)
            divisions[d].weights[row] = np.frombuffer(body, "<f4", C, 12).copy()
        elif tag == b"PSTA":
            d, C, mat_len, dummy, uncompressed, freq_len = struct.unpack_from("<6i", body, 0)
            tab = np.frombuffer(body, "<i4", 3 * C, 24).copy().reshape(3, C)
            divisions[d].pattern_states = dict(state_counts=tab[0], matrix_offsets=tab[1], freq_offsets=tab[2],
                                               matrix_length=mat_len, dummy_patterns=dummy, uncompressed_sites=uncompressed)
        elif tag == b"EIGN":
            d, eig, S = struct.unpack_from("<3i", body, 0)
            a = np.frombuffer(body, "<f8", S + 2 * S * S, 12).copy()
            parts, part, eig = (eig >> 24) & 0xff, (eig >> 16) & 0xff, eig & 0xffff
            lam, V, Vinv = a[:S], a[S:S + S * S].reshape(S, S), a[S + S * S:].reshape(S, S)
            if parts <= 1:
                events.append(Event("eigen", d, eig, lam=lam, V=V, Vinv=Vinv))
            else:
                # one eigensystem per category (NY98): rebuild the slot's cijk block, part by part, the way
                # CalcCijk does (src/utils.c:9734-9746: c[i][j][k] = u[i][k] * v[k][j]; IEEE products)
                blk = np.concatenate([lam, np.zeros(S), (V[:, None, :] * Vinv.T[None, :, :]).ravel()])
                if part == 0:
                    pending_parts = []
                pending_parts.append(blk)
                if part == parts - 1:
                    events.append(Event("cijk", d, eig, block=np.concatenate(pending_parts)))
        elif tag == b"CIJK":
            d, eig, S = struct.unpack_from("<3i", body, 0)
            a = np.frombuffer(body, "<f8", (len(body) - 12) // 8, 12).copy()      # all parts of the slot
            if eig == abi.EIGEN_INLINE:
                pending_inline = a                # travels with the next evaluation
            else:
                events.append(Event("cijk", d, eig, block=a))
        elif tag == b"EVAL":
            (d, chain, nmat, nop, sdst, ssrc, root, wrow, flags, haspi, K, S) = struct.unpack_from("<12i", body, 0)
            off = 48
            (pinv,) = struct.unpack_from("<d", body, off); off += 8
            rates = np.frombuffer(body, "<f8", K, off).copy(); off += 8 * K
            catw = np.frombuffer(body, "<f8", K, off).copy(); off += 8 * K
            freqs = np.frombuffer(body, "<f8", S, off).copy(); off += 8 * S
            mats = np.frombuffer(body, abi.MAT_DTYPE, nmat, off).copy(); off += 16 * nmat
            ops = np.frombuffer(body, abi.OP_DTYPE, nop, off).copy(); off += 36 * nop
            (lnl,) = struct.unpack_from("<d", body, off); off += 8
            (aborted, _pad) = struct.unpack_from("<2i", body, off)
            spec = abi.EvalSpec(mats=mats, ops=ops, site_dst=sdst, site_src=ssrc, root=root, weights_row=wrow,
                                flags=flags, p_invar=pinv, has_p_invar=haspi, rates=rates, cat_weights=catw,
                                freqs=freqs, chain=chain, division=d, lnl_ref=lnl, aborted=aborted,
                                inline_eigen=pending_inline)
            pending_inline = None
            events.append(Event("eval", d, spec=spec))
        else:
            raise ValueError(f"{path}: unknown chunk {tag!r}")
    return divisions, events


def make_instance(lib: abi.Library, div: Division, device: int = 0, max_evaluations: int = 1) -> abi.Instance:
    """Create an instance for a recorded division and load its tips and weights."""
    c = dict(div.cfg)
    c["device"] = device
    c["max_evaluations"] = max(max_evaluations, 1)
    inst = abi.Instance(lib, **c)
    if div.pattern_states is not None:
        inst.set_pattern_states(**div.pattern_states)
    for tip, m in div.tips.items():
        inst.set_tip_states(tip, m)
    for row, w in div.weights.items():
        inst.set_pattern_weights(row, w)
    return inst


def apply_event(inst: abi.Instance, ev: Event):
    """Apply a non-eval event (eigensystem upload) to an instance."""
    if ev.kind == "eigen":
        inst.set_eigen_decomposition(ev.eigen, ev.V, ev.Vinv, ev.lam)
    elif ev.kind == "cijk":
        inst.set_cijk(ev.eigen, ev.block)
    else:
        raise ValueError(ev.kind)


def replay(lib: abi.Library, path, max_evals: int | None = None, device: int = 0, arith: int | None = None):
    """Replay a record file; -> list of (spec, lnL, status) in evaluation order."""
    divisions, events = load(path)
    insts = {d: make_instance(lib, div, device) for d, div in divisions.items()}
    if arith is not None:
        for inst in insts.values():
            inst.set_arith(arith)
    out = []
    try:
        for ev in events:
            inst = insts[ev.division]
            if ev.kind == "eval":
                lnl, st = inst.evaluate(ev.spec)
                out.append((ev.spec, float(lnl[0]), int(st[0])))
                if max_evals is not None and len(out) >= max_evals:
                    break
            else:
                apply_event(inst, ev)
    finally:
        for inst in insts.values():
            inst.close()
    return out
