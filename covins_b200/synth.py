"""Seeded synthetic inputs for the hot path — numpy only, deterministic.

The reference ships no fixtures, maps or tests, so every parity/bench input is
generated here.  Sizes follow BASELINE.json configs: C1 200 KF/10k LM, C2 800/40k, C3 2000/100k
(5 agents), C5 10000/1M (12 agents); 1000 ORB features per KF (32 B each,
covins_frontend/config/EuRoC.yaml:52-53), 300 SIFT (128 x f32).
"""
from __future__ import annotations

import numpy as np

ORB_BYTES = 32
SIFT_DIM = 128


def _flip_mask(rng: np.random.Generator, shape) -> np.ndarray:
    """Random byte mask with each bit set w.p. 1/16 (= AND of 4 uniform bytes) ~ the 0.06 of §8d."""
    m = rng.integers(0, 256, shape, dtype=np.uint8)
    for _ in range(3):
        m &= rng.integers(0, 256, shape, dtype=np.uint8)
    return m


def orb_keyframes(seed: int, n_kf: int, n_feat: int = 1000, lm_frac: float = 0.4, n_lm: int | None = None,
                  window: int = 4000):
    """ORB-like descriptor sets for n_kf keyframes.

    Each landmark owns a random 256-bit code; an observation is that code with bits flipped w.p.
    1/16 (matched Hamming ~ 16 +- 4, unmatched ~ 128 +- 8).  Keypoints without a landmark get a
    random code.  Landmark ids seen by keyframe i are drawn from a sliding window of the pool so
    that neighbouring keyframes are covisible.

    Returns desc [n_kf, n_feat, 32] u8 and lm_id [n_kf, n_feat] i32 (-1 = keypoint has no landmark;
    the reference's skip mask, LandmarkMatchingAlgorithm.cpp:76-84).
    """
    rng = np.random.default_rng(seed)
    n_with = int(round(n_feat * lm_frac))
    if n_lm is None:
        n_lm = max(window, n_kf * n_with // 8)
    window = min(window, n_lm)
    codes = rng.integers(0, 256, (n_lm, ORB_BYTES), dtype=np.uint8)
    desc = rng.integers(0, 256, (n_kf, n_feat, ORB_BYTES), dtype=np.uint8)
    lm_id = np.full((n_kf, n_feat), -1, np.int32)
    for i in range(n_kf):
        lo = 0 if n_kf == 1 else int((n_lm - window) * i / (n_kf - 1))
        ids = lo + rng.choice(window, size=n_with, replace=False)
        slots = rng.choice(n_feat, size=n_with, replace=False)
        desc[i, slots] = codes[ids] ^ _flip_mask(rng, (n_with, ORB_BYTES))
        lm_id[i, slots] = ids
    return desc, lm_id


def sift_keyframes(seed: int, n_kf: int, n_feat: int = 300, lm_frac: float = 0.4, n_lm: int | None = None,
                   window: int = 2000, noise: float = 6.0):
    """SIFT-like descriptors: integer-valued float32 in [0,255] like cv::xfeatures2d::SIFT output
    (covins_frontend/src/frontend_wrapper.cpp:603): landmark code ~ clipped Gamma, L2-normalised to
    512 then clipped; observation = code + round(N(0, noise)), clipped."""
    rng = np.random.default_rng(seed)
    n_with = int(round(n_feat * lm_frac))
    if n_lm is None:
        n_lm = max(window, n_kf * n_with // 8)
    window = min(window, n_lm)

    def codes_(n):
        g = rng.gamma(0.7, 1.0, (n, SIFT_DIM))
        g = g / np.linalg.norm(g, axis=1, keepdims=True) * 512.0
        return np.clip(np.rint(g), 0, 255)

    codes = codes_(n_lm)
    desc = codes_(n_kf * n_feat).reshape(n_kf, n_feat, SIFT_DIM)
    lm_id = np.full((n_kf, n_feat), -1, np.int32)
    for i in range(n_kf):
        lo = 0 if n_kf == 1 else int((n_lm - window) * i / (n_kf - 1))
        ids = lo + rng.choice(window, size=n_with, replace=False)
        slots = rng.choice(n_feat, size=n_with, replace=False)
        desc[i, slots] = np.clip(codes[ids] + np.rint(rng.normal(0, noise, (n_with, SIFT_DIM))), 0, 255)
        lm_id[i, slots] = ids
    return desc.astype(np.float32), lm_id


def seg_ptr_uniform(n_seg: int, n_per: int) -> np.ndarray:
    return (np.arange(n_seg + 1, dtype=np.int64) * n_per).astype(np.int32)
