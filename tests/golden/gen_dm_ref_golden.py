"""Generates tests/golden/dense_matcher_ref.npz with the REFERENCE's own estd2::DenseMatcher, compiled from
/root/reference/covins_backend/src/dense_matcher/*.cpp + src/matcher/MatchingAlgorithm.cpp by oracle/ref/Makefile
(oracle/_ref/libdm_ref.so, driver oracle/ref/dm_ref_shim.cpp) — reference-produced vectors for SURVEY §8a M5/M6.

Run once in the build container:  python tests/golden/gen_dm_ref_golden.py
The fixture is committed; the tests need neither /root/reference nor oracle/_ref at run time.
All cases use numMatcherThreads = 1 (the canonical deterministic order); one C2-sized case is also run
with 8 threads as in placerec_be.cpp:87 and stored for information (it coincides on tie-free data).
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from covins_b200 import synth  # noqa: E402
from oracle import ref_dm  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def tie_case(seed, nA, nB, n_codes, p_skip):
    """few distinct codes + tiny perturbations → many equal distances, steals and displacement chains"""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (n_codes, 32), dtype=np.uint8)

    def mk(n):
        d = base[rng.integers(0, n_codes, n)].copy()
        return d ^ (rng.integers(0, 256, (n, 32), dtype=np.uint8) & rng.integers(0, 256, (n, 32), dtype=np.uint8)
                    & rng.integers(0, 256, (n, 32), dtype=np.uint8))
    return mk(nA), (rng.random(nA) < p_skip).astype(np.uint8), mk(nB), (rng.random(nB) < p_skip).astype(np.uint8)


def main():
    cases = {}

    def add(name, A, sA, B, sB, thr=50.0, nb=4, threads=1):
        a, b, d = ref_dm.dense_match(A, sA, B, sB, thr, threads, nb)
        cases[name] = dict(A=A, skipA=sA, B=B, skipB=sB, thr=np.float32(thr), num_best=np.int32(nb), outA=a, outB=b, outD=d)
        print(f"{name}: nA={len(A)} nB={len(B)} num_best={nb} -> {len(a)} matches")

    for s in range(3):
        add(f"ties_{s}", *tie_case(100 + s, 200, 260, 10, 0.2))
    for nb in (1, 2, 3):
        add(f"ties_numbest{nb}", *tie_case(7, 150, 170, 8, 0.1), nb=nb)
    add("ties_noskip", *[x if i % 2 == 0 else np.zeros_like(x) for i, x in enumerate(tie_case(9, 120, 90, 6, 0.0))])
    # C2-sized keyframe pairs (1000 features, 40 % with landmarks, covisible): placerec_be.cpp:85-90
    desc, lm = synth.orb_keyframes(seed=42, n_kf=4, n_feat=1000, n_lm=1200, window=1000)
    for s in range(3):
        add(f"c2_pair_{s}", desc[0], (lm[0] < 0).astype(np.uint8), desc[s + 1], (lm[s + 1] < 0).astype(np.uint8))
    add("c2_pair_0_threads8", desc[0], (lm[0] < 0).astype(np.uint8), desc[1], (lm[1] < 0).astype(np.uint8), threads=8)
    # degenerate shapes: empty B, single rows, everything skipped
    e = np.zeros((0, 32), np.uint8); z = np.zeros(0, np.uint8)
    add("empty_B", desc[0][:50], np.zeros(50, np.uint8), e, z)
    add("single", desc[0][:1], np.zeros(1, np.uint8), desc[0][:1], np.zeros(1, np.uint8))
    add("all_skipped", desc[0][:64], np.ones(64, np.uint8), desc[1][:64], np.zeros(64, np.uint8))
    flat = {f"{n}/{k}": v for n, c in cases.items() for k, v in c.items()}
    np.savez_compressed(os.path.join(HERE, "dense_matcher_ref.npz"), **flat)


if __name__ == "__main__":
    main()
