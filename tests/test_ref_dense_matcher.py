"""M5/M6 parity pinned by REFERENCE-PRODUCED vectors: tests/golden/dense_matcher_ref.npz was written by the reference's
own estd2::DenseMatcher (oracle/_ref/libdm_ref.so, compiled from /root/reference by oracle/ref/Makefile; generator
tests/golden/gen_dm_ref_golden.py).  CPU: the oracle restatement must reproduce every case; when oracle/_ref is
present (build container) it is re-run live as well.  GPU: the CUDA path (through the C-ABI) must reproduce them."""
import os
import numpy as np
import pytest

from oracle import knn as ora, ref_dm
from conftest import golden_cases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dense_matcher_ref.npz")


def _cases():
    g, names = golden_cases(GOLD)
    for n in names:
        yield n, {k: g[f"{n}/{k}"] for k in ("A", "skipA", "B", "skipB", "thr", "num_best", "outA", "outB", "outD")}


def test_oracle_reproduces_reference_dense_matcher():
    n_cases = 0
    for name, c in _cases():
        a, b, d = ora.landmark_match(c["A"], c["skipA"], c["B"], c["skipB"], float(c["thr"]), int(c["num_best"]))
        assert np.array_equal(a, c["outA"]) and np.array_equal(b, c["outB"]) and np.array_equal(d, c["outD"]), name
        n_cases += 1
    assert n_cases >= 14


@pytest.mark.skipif(not ref_dm.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_live_reference_equals_fixture_and_oracle():
    """the fixture is what the reference code produces today; 8 matcher threads (placerec_be.cpp:87) give the same result
    on the tie-free C2 keyframe pairs"""
    for name, c in _cases():
        a, b, d = ref_dm.dense_match(c["A"], c["skipA"], c["B"], c["skipB"], float(c["thr"]), 1, int(c["num_best"]))
        assert np.array_equal(a, c["outA"]) and np.array_equal(b, c["outB"]) and np.array_equal(d, c["outD"]), name
    rng = np.random.default_rng(77)
    for trial in range(20):   # random shapes, heavy ties: reference (1 thread) == oracle
        nA, nB = int(rng.integers(1, 90)), int(rng.integers(1, 90))
        base = rng.integers(0, 256, (5, 32), dtype=np.uint8)
        mk = lambda n: base[rng.integers(0, 5, n)] ^ (rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8))
        A, B = mk(nA), mk(nB)
        sA = (rng.random(nA) < 0.25).astype(np.uint8); sB = (rng.random(nB) < 0.25).astype(np.uint8)
        nb = int(rng.integers(1, 5))
        r = ref_dm.dense_match(A, sA, B, sB, 50.0, 1, nb)
        o = ora.landmark_match(A, sA, B, sB, 50.0, nb)
        assert all(np.array_equal(x, y) for x, y in zip(r, o)), trial


@pytest.mark.gpu
def test_cuda_reproduces_reference_dense_matcher(ctx):
    from covins_b200 import matching as M
    for name, c in _cases():
        (a, b, d), = M.landmark_match(ctx, c["A"], c["skipA"], c["B"], c["skipB"], None, float(c["thr"]), int(c["num_best"]))
        assert np.array_equal(a, c["outA"]) and np.array_equal(b, c["outB"]) and np.array_equal(d, c["outD"]), name
    # all C2 pairs in one batched launch (one segment per candidate keyframe, placerec_be.cpp:75-112)
    cs = {n: c for n, c in _cases() if n.startswith("c2_pair_") and "threads" not in n}
    names = sorted(cs)
    B = np.concatenate([cs[n]["B"] for n in names]); sB = np.concatenate([cs[n]["skipB"] for n in names])
    seg = np.concatenate([[0], np.cumsum([len(cs[n]["B"]) for n in names])]).astype(np.int32)
    out = M.landmark_match(ctx, cs[names[0]]["A"], cs[names[0]]["skipA"], B, sB, seg, 50.0, 4)
    for s, n in enumerate(names):
        assert np.array_equal(out[s][0], cs[n]["outA"]) and np.array_equal(out[s][1], cs[n]["outB"]) and np.array_equal(out[s][2], cs[n]["outD"]), n
