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


def _lm_desc_case(seed, sizes):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (7, 32), dtype=np.uint8)
    rows = []
    for n in sizes:
        d = base[rng.integers(0, 7, n)].copy()
        flip = rng.random(d.shape) < 0.08
        d[flip] ^= rng.integers(1, 256, int(flip.sum()), dtype=np.uint8)
        rows.append(d)
    lm_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    return (np.concatenate(rows) if sum(sizes) else np.zeros((0, 32), np.uint8)), lm_ptr


def test_landmark_descriptor_oracle_vs_independent_restatement():
    """Landmark::ComputeDescriptor (landmark_be.cpp:49-92): C oracle vs a plain numpy restatement (full distance matrix,
    per-row sort, median index (int)(0.5 (n-1)), first strict minimum), incl. empty / single / duplicated observers."""
    sizes = [0, 1, 2, 3, 8, 8, 33, 5, 0, 64, 2]
    cand, lm_ptr = _lm_desc_case(3, sizes)
    cand[lm_ptr[4]:lm_ptr[4] + 8] = cand[lm_ptr[4]]            # all observers identical → row 0 wins
    best, desc = ora.landmark_descriptor(cand, lm_ptr)
    bits = np.unpackbits(cand, axis=1).astype(np.int32)
    for l, n in enumerate(sizes):
        if n == 0:
            assert best[l] == -1
            continue
        B = bits[lm_ptr[l]:lm_ptr[l + 1]]
        D = (B[:, None, :] != B[None, :, :]).sum(-1).astype(np.float64)
        med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        ref = int(np.argmin(med))                              # argmin returns the first minimum
        assert best[l] == ref, (l, n)
        assert np.array_equal(desc[l], cand[lm_ptr[l] + ref])
    assert best[4] == 0
