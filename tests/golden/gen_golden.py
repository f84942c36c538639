"""Generates tests/golden/knn_*.npz with the ONE executable piece of the reference's matching path that
exists in the build container: OpenCV's BFMatcher (python cv2 4.13) — the library call the reference
makes at covins_backend/src/covins_backend/placerec_gen_be.cpp:99 and RelNonCentralPosSolver.cpp:323.

Run once in the build container:  python tests/golden/gen_golden.py
The fixtures are committed; neither the tests nor bench.py need cv2 or /root/reference at run time.
"""
import os
import sys
import numpy as np
import cv2

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from covins_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def bf_knn(q, t, norm, k=2):
    res = cv2.BFMatcher(norm).knnMatch(q, t, k=k)
    idx = np.full((len(q), k), -1, np.int32)
    dist = np.full((len(q), k), np.inf, np.float32)
    for r in res:
        for c, m in enumerate(r):
            idx[m.queryIdx, c] = m.trainIdx
            dist[m.queryIdx, c] = m.distance
    return idx, dist


def ref_filter(idx, dist, thr, ratio):
    """placerec_gen_be.cpp:102-114 in numpy float32."""
    thr = np.float32(thr); ratio = np.float32(ratio)
    ok = (idx[:, 0] >= 0) & (idx[:, 1] >= 0) & (dist[:, 0] <= thr) & (dist[:, 0] < ratio * dist[:, 1])
    return np.where(ok, idx[:, 0], -1).astype(np.int32)


def main():
    cases = {}
    rng = np.random.default_rng(1234)
    # 1. ORB keyframe pair with shared landmarks (the per-candidate call of placerec_gen_be.cpp:99)
    d, _ = synth.orb_keyframes(seed=1, n_kf=2, n_feat=300, n_lm=200, window=200)
    cases["orb_pair"] = (d[0], d[1])
    # 2. heavy ties: train rows duplicated, few distinct codes
    base = rng.integers(0, 256, (6, 32), dtype=np.uint8)
    t = base[rng.integers(0, 6, 64)]
    q = base[rng.integers(0, 6, 16)] ^ (rng.integers(0, 256, (16, 32), dtype=np.uint8) & 1)
    cases["orb_ties"] = (q, t)
    # 3. tiny train sets (nt = 2 and nt = 1 < k) and a single query
    cases["orb_nt2"] = (rng.integers(0, 256, (5, 32), dtype=np.uint8), rng.integers(0, 256, (2, 32), dtype=np.uint8))
    cases["orb_nt1"] = (rng.integers(0, 256, (5, 32), dtype=np.uint8), rng.integers(0, 256, (1, 32), dtype=np.uint8))
    cases["orb_nq1"] = (rng.integers(0, 256, (1, 32), dtype=np.uint8), rng.integers(0, 256, (37, 32), dtype=np.uint8))
    # 4. all-zero vs all-one (distance 256) and identical (distance 0)
    q = np.zeros((3, 32), np.uint8)
    t = np.concatenate([np.full((2, 32), 255, np.uint8), np.zeros((2, 32), np.uint8)])
    cases["orb_extreme"] = (q, t)
    # 5. ragged sizes not multiples of any tile
    cases["orb_ragged"] = (rng.integers(0, 256, (131, 32), dtype=np.uint8),
                           rng.integers(0, 256, (1027, 32), dtype=np.uint8))
    out = {}
    for name, (q, t) in cases.items():
        idx, dist = bf_knn(q, t, cv2.NORM_HAMMING)
        out[name + "/q"] = q; out[name + "/t"] = t; out[name + "/idx"] = idx; out[name + "/dist"] = dist
        out[name + "/match"] = ref_filter(idx, dist, 40.0, 0.8)
    np.savez_compressed(os.path.join(HERE, "knn_hamming.npz"), **out)

    out = {}
    s, _ = synth.sift_keyframes(seed=2, n_kf=2, n_feat=150, n_lm=120, window=120)
    cases = {"sift_pair": (s[0], s[1])}
    t = s[1][rng.integers(0, 10, 40)]  # duplicated rows → ties
    cases["sift_ties"] = (s[0][:20], t)
    cases["sift_nt2"] = (s[0][:5], s[1][:2])
    cases["sift_ragged"] = (s[0][:77], np.concatenate([s[1], s[0][77:]])[:211])
    for name, (q, t) in cases.items():
        idx, dist = bf_knn(q, t, cv2.NORM_L2)
        out[name + "/q"] = q.astype(np.uint8); out[name + "/t"] = t.astype(np.uint8)  # integer-valued → compact
        out[name + "/idx"] = idx; out[name + "/dist"] = dist
        out[name + "/match"] = ref_filter(idx, dist, 500.0, 0.8)
    np.savez_compressed(os.path.join(HERE, "knn_l2.npz"), **out)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
