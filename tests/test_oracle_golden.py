"""CPU: the oracle restatement against the reference's own numbers.

Every golden file is a recording of the UNMODIFIED reference (tests/golden/make_golden.sh);
the oracle must reproduce each recorded lnL essentially bit for bit (<= 1e-12 relative,
observed 0), which pins it before any CUDA result is compared with it."""
import numpy as np
import pytest

from conftest import GOLDEN, GOLDEN_FILES
from mrbayes_b200 import abi, records


@pytest.mark.parametrize("stem,arith", GOLDEN_FILES)
def test_oracle_matches_reference(oracle_lib, stem, arith):
    res = records.replay(oracle_lib, GOLDEN / f"{stem}.gold.gz", arith=arith)
    assert len(res) >= 40
    rel = np.array([abs(l - s.lnl_ref) / abs(s.lnl_ref) for s, l, _ in res])
    assert all(st == abi.EVAL_OK for _, _, st in res)
    assert rel.max() <= 1e-12, f"{stem}: oracle drifted from the reference ({rel.max():.3e})"


def test_known_answers_primates(oracle_lib):
    """SURVEY 8c anchors: initial lnL of the 8 chains of primates GTR+G4, seed 12345."""
    res = records.replay(oracle_lib, GOLDEN / "primates_gtr_g4_fma.gold.gz", max_evals=8, arith=1)
    want = [-8955.8613832833762, -8351.2408433520559, -8692.296794860029, -8803.6253014443209]
    for (spec, lnl, _), w in zip(res[:4], want):
        assert lnl == pytest.approx(w, rel=1e-15)
    want2 = [-8949.005293, -8654.647788, -8816.627216, -8175.682811]
    for (spec, lnl, _), w in zip(res[4:8], want2):
        assert lnl == pytest.approx(w, abs=1e-6)


def test_known_answers_other_datasets(oracle_lib):
    r = records.replay(oracle_lib, GOLDEN / "ovomucoids_wag_g4_sse.gold.gz", max_evals=1, arith=0)
    assert r[0][1] == pytest.approx(-6831.225867, abs=1e-6)
    r = records.replay(oracle_lib, GOLDEN / "replicase_m0_sse.gold.gz", max_evals=1, arith=0)
    assert r[0][1] == pytest.approx(-8974.126819, abs=1e-6)


def test_sse_and_fma_variants_differ_only_in_rounding(oracle_lib):
    """Calibration of the 1e-6 tolerance: the reference's own SSE and FMA builds differ by
    ~1e-9 relative on the same state ('Trajectory vs evaluation parity')."""
    a = records.replay(oracle_lib, GOLDEN / "primates_gtr_g4_fma.gold.gz", max_evals=8, arith=1)
    b = records.replay(oracle_lib, GOLDEN / "primates_gtr_g4_fma.gold.gz", max_evals=8, arith=0)
    rel = max(abs(x[1] - y[1]) / abs(x[1]) for x, y in zip(a, b))
    assert 0 < rel < 1e-7


def test_record_file_structure():
    divs, events = records.load(GOLDEN / "primates_gtr_g4_fma.gold.gz")
    assert list(divs) == [0]
    d = divs[0]
    assert d.cfg["pattern_count"] == 413 and d.cfg["state_count"] == 4 and d.cfg["category_count"] == 4
    assert d.cfg["tip_count"] == 12 and len(d.tips) == 12
    assert d.cfg["partials_count"] == 12 + 9 * 10          # (chains+1)*nIntNodes + tips
    w = d.weights[0]
    assert w.sum() == 898                                   # sites of primates.nex
    evals = [e for e in events if e.kind == "eval"]
    assert len(evals) == 400
    # the first evaluation of every chain is a full one: 21 matrices, 10 node updates
    assert len(evals[0].spec.mats) == 21 and len(evals[0].spec.ops) == 10
    assert evals[0].spec.site_src == abi.NONE
