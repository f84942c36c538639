"""CPU: the oracle restatement vs the committed cv2.BFMatcher golden vectors (tests/golden/gen_golden.py),
plus DenseMatcher-restatement invariants.  No GPU needed."""
import os

import numpy as np
import pytest

from conftest import golden_cases
from oracle import knn as ora


def test_hamming_oracle_matches_cv2_golden(golden_dir):
    g, names = golden_cases(os.path.join(golden_dir, "knn_hamming.npz"))
    assert len(names) >= 7
    for n in names:
        idx, dist = ora.knn_hamming(g[n + "/q"], g[n + "/t"], k=2)
        assert np.array_equal(idx, g[n + "/idx"]), n
        d = np.where(idx >= 0, dist.astype(np.float32), np.inf)
        assert np.array_equal(d, g[n + "/dist"]), n
        mt, _, cnt = ora.ratio_filter(idx, d, 40.0, 0.8)
        assert np.array_equal(mt, g[n + "/match"]), n
        assert cnt == (g[n + "/match"] >= 0).sum()


def test_l2_oracle_matches_cv2_golden(golden_dir):
    g, names = golden_cases(os.path.join(golden_dir, "knn_l2.npz"))
    for n in names:
        idx, dist = ora.knn_l2(g[n + "/q"].astype(np.float32), g[n + "/t"].astype(np.float32), k=2)
        assert np.array_equal(idx, g[n + "/idx"]), n
        d = np.where(idx >= 0, dist, np.inf)
        assert np.array_equal(d, g[n + "/dist"]), n  # bit-exact: integer-valued SIFT
        mt, _, _ = ora.ratio_filter(idx, d, 500.0, 0.8)
        assert np.array_equal(mt, g[n + "/match"]), n


def test_hamming256_bit_hack_equals_popcount():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (64, 32), dtype=np.uint8)
    for i in range(64):
        assert ora.hamming256(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())


def test_batch_equals_per_segment_calls():
    rng = np.random.default_rng(3)
    q = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    seg = np.array([0, 100, 100, 157, 300], np.int32)  # includes an empty segment
    idx, dist = ora.knn_hamming_batch(q, t, seg, k=2)
    for s in range(4):
        i1, d1 = ora.knn_hamming(q, t[seg[s]:seg[s + 1]], k=2)
        assert np.array_equal(idx[s], i1) and np.array_equal(dist[s], d1)


def _py_dense_matcher(A, skipA, B, skipB, thr=50.0, nb=4):
    """Independent pure-python restatement (small sizes) of DenseMatcher.hpp:152-220 + DenseMatcher.cpp:62-104."""
    nA, nB = len(A), len(B)
    FM = np.finfo(np.float32).max
    best = [[(-1, thr)] * nb for _ in range(nA)]
    vp = [(-1, FM)] * nB

    def dist(a, b):
        d = float(np.unpackbits(A[a] ^ B[b]).sum())
        return d if d < thr else FM

    def assign(a, start):
        for k in range(start, nb):
            b, d = best[a][k]
            if b == -1:
                return
            if vp[b][0] == -1:
                vp[b] = (a, d)
                return
            if d < vp[b][1]:
                old = vp[b][0]
                vp[b] = (a, d)
                assign(old, 1)
                return

    for a in range(nA):
        if skipA[a]:
            continue
        lst = best[a]
        for b in range(nB):
            if skipB[b]:
                continue
            d = dist(a, b)
            if d < lst[nb - 1][1]:
                lb = 0
                while lb < nb and lst[lb][1] < d:
                    lb += 1
                lst.insert(lb, (b, d))
                lst.pop()
        assign(a, 0)
    return [(vp[b][0], b, vp[b][1]) for b in range(nB) if vp[b][1] < thr]


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_landmark_match_oracle_vs_python_restatement(seed):
    rng = np.random.default_rng(seed)
    # few distinct codes + tiny perturbations → many ties and displacement chains
    base = rng.integers(0, 256, (12, 32), dtype=np.uint8)
    def mk(n):
        d = base[rng.integers(0, 12, n)].copy()
        flips = rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8) \
            & rng.integers(0, 256, (n, 32), dtype=np.uint8)
        return d ^ flips
    A, B = mk(60), mk(70)
    skipA = (rng.random(60) < 0.3).astype(np.uint8); skipB = (rng.random(70) < 0.3).astype(np.uint8)
    oa, ob, od = ora.landmark_match(A, skipA, B, skipB, thr=50.0, num_best=4)
    ref = _py_dense_matcher(A, skipA, B, skipB)
    assert [(int(a), int(b), float(d)) for a, b, d in zip(oa, ob, od)] == [(a, b, float(d)) for a, b, d in ref]
    assert len(ref) > 5
    # invariants: one-to-one, no skipped keypoint is ever matched, every distance < 50
    assert len(set(oa.tolist())) == len(oa) and len(set(ob.tolist())) == len(ob)
    assert not skipA[oa].any() and not skipB[ob].any() and (od < 50).all()
