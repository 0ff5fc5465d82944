"""The seam (mrbayes_b200/seam/mb200_seam.c) inside the reference's own `mb` program.

oracle/_ref/mb_b200 = the UNMODIFIED reference objects + the seam + oracle/ref_harness.c (ld --wrap on
the single call site of the hot path).  Two kinds of tests:

* CPU (no GPU): in `dump` mode the seam's engine calls are recorded, not executed.  The records written
  through the node-granular function-pointer forms (TiProbs_B200 ... Likelihood_B200 installed in
  ModelInfo, driven by the reference's own LaunchLogLikeForDivision loop) must be byte-identical to the
  ones written through the seam's own loop -- and to the committed golden files, which pins the flip
  order, the operation lists, the rate assembly and the scaler bookkeeping of both paths.
* GPU: `shadow` mode -- the reference's CPU kernels drive a real MCMC run while every evaluation also
  runs on the engine; |lnL_gpu - lnL_cpu| / |lnL_cpu| is checked per evaluation (bar: 1e-6, the
  north-star tolerance; observed ~1e-8).  `gpu` mode: the engine alone drives the chain, through the
  seam's loop and through the function-pointer forms: identical lnL streams.
"""
from __future__ import annotations

import gzip
import json
import os
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "oracle" / "_ref" / "mb_b200"
CMD = ROOT / "tests" / "golden" / "cmd"

needs_harness = pytest.mark.skipif(not BIN.exists(), reason="oracle/_ref/mb_b200 not built (needs /root/reference at build time)")


def run_harness(tmp_path: Path, stem: str, ngen: int, mode: str, via: str = "seam", extra_env=None, timeout=900):
    nex = tmp_path / f"{stem}.{mode}.{via}.nex"
    text = (CMD / f"{stem}.nex").read_text().replace("NGEN", str(ngen)).replace("OUTPREFIX", str(tmp_path / f"out_{stem}_{mode}_{via}"))
    nex.write_text(text)
    report = tmp_path / f"{stem}.{mode}.{via}.json"
    env = dict(os.environ, MB200_MODE=mode, MB200_REPORT=str(report), MB200_VIA=via)
    env.update(extra_env or {})
    p = subprocess.run([str(BIN), str(nex)], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert report.exists(), p.stdout[-2000:] + p.stderr[-2000:]
    rep = json.loads(report.read_text().strip().splitlines()[-1])
    rep["stderr"] = p.stderr[-2000:]
    return rep


# (command stem, generations, max evaluations recorded, committed golden written from the same command or None)
RECORD_CASES = [
    ("primates_gtr_g4", 60, 400, "primates_gtr_g4_fma"),
    ("primates_hky_g4", 100, 200, "primates_hky_g4_fma"),
    ("primates_f81_i", 100, 150, "primates_f81_i_fma"),
    ("cynmix_full", 40, 200, "cynmix_full_fma"),
    ("replicase_ny98", 60, 40, "replicase_ny98_sse"),
]


@needs_harness
@pytest.mark.parametrize("stem,ngen,max_evals,golden", RECORD_CASES)
def test_function_pointer_forms_record_the_same_evaluations(tmp_path, stem, ngen, max_evals, golden):
    files = {}
    for via in ("seam", "fnptr"):
        out = tmp_path / f"{stem}.{via}.gold"
        rep = run_harness(tmp_path, stem, ngen, "dump", via, {"MB200_DUMP_FILE": str(out), "MB200_DUMP_MAX": str(max_evals)})
        assert rep["dumped"] > 0 and rep["via"] == via
        files[via] = out.read_bytes()
    assert files["seam"] == files["fnptr"], "function-pointer forms and the seam's own loop disagree on the evaluation records"
    if golden is not None:
        committed = gzip.open(ROOT / "tests" / "golden" / f"{golden}.gold.gz").read()
        assert files["seam"] == committed, f"seam no longer reproduces tests/golden/{golden}.gold.gz"


# ---------------------------------------------------------------------------------------- GPU
SHADOW_CASES = [
    # stem, generations, expected unsupported calls (None = any), min evaluations
    ("primates_gtr_g4", 2000, 0, 16000),
    ("primates_gtr_ig4", 500, 0, 1000),
    ("primates_hky_g4", 500, 0, 1000),
    ("primates_f81_i", 300, 0, 500),
    ("ovomucoids_wag_g4", 300, 0, 300),
    ("replicase_m0", 200, 0, 200),
    ("replicase_ny98", 100, 0, 100),
    ("cynmix_full", 300, 0, 1500),
]


@needs_harness
@pytest.mark.gpu
@pytest.mark.parametrize("stem,ngen,unsupported,min_calls", SHADOW_CASES)
def test_shadow_run_matches_reference_per_evaluation(tmp_path, engine_lib, stem, ngen, unsupported, min_calls):
    rep = run_harness(tmp_path, stem, ngen, "shadow")
    assert rep["calls"] >= min_calls, rep
    assert rep["failed"] == 0, rep
    assert rep["compared"] == rep["calls"] - rep["unsupported_calls"], rep
    if unsupported is not None:
        assert rep["unsupported_calls"] == unsupported, rep
    assert rep["max_rel"] < 1e-6, rep          # north-star tolerance: 1e-6 relative per evaluation


@needs_harness
@pytest.mark.gpu
@pytest.mark.parametrize("stem,ngen", [("primates_gtr_g4", 1000), ("cynmix_full", 100), ("replicase_m0", 60)])
def test_engine_drives_the_chain_identically_through_both_entry_levels(tmp_path, engine_lib, stem, ngen):
    a = run_harness(tmp_path, stem, ngen, "gpu", "seam", {"MB200_MULTIPART": "0"})
    b = run_harness(tmp_path, stem, ngen, "gpu", "fnptr")
    assert a["calls"] == b["calls"] and a["calls"] > 0
    assert a["unsupported_calls"] == 0 and b["unsupported_calls"] == 0
    assert a["lnl_hash"] == b["lnl_hash"], (a, b)       # the same lnL, bit for bit, at every evaluation
    c = run_harness(tmp_path, stem, ngen, "gpu", "seam")    # partition-batched LogLike (MB200LogLike)
    assert c["unsupported_calls"] == 0 and c["aborts"] == a["aborts"]
