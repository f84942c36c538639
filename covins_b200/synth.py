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


def _rot(rvec):
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def se3_search_scene(seed: int, n_kp: int = 1000, n_shared: int = 300, pix_noise: float = 1.0, pose_noise: float = 0.01,
                     img_w: int = 752, img_h: int = 480, n_octaves_data: int = 1):
    """Two keyframes looking at a common set of landmarks, as FeatureMatcher::SearchBySE3 sees them after the first RANSAC
    (feature_matcher_be.cpp:293-498): per keyframe keypoints (float), octaves, ORB descriptors, and for the keypoints that
    carry a landmark its world position / max distance / representative descriptor.  KF1 and KF2 each own DIFFERENT landmark
    objects for the same physical points (that is what the search is meant to associate).  Returns plain dict inputs for
    covins_b200.placerec.KfView plus the relative pose T12 (slightly perturbed) and the already-matched masks."""
    rng = np.random.default_rng(seed)
    K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1.0]])
    # camera 1 at the origin, camera 2 displaced; points in front of both
    Rwc = [np.eye(3), _rot(rng.normal(0, 0.08, 3))]
    twc = [np.zeros(3), rng.normal(0, 0.25, 3)]
    n_pts = 3 * n_shared
    uv = np.stack([rng.uniform(20, img_w - 20, n_pts), rng.uniform(20, img_h - 20, n_pts)], -1)
    depth = rng.uniform(2.0, 12.0, n_pts)
    rays = np.linalg.solve(K, np.concatenate([uv, np.ones((n_pts, 1))], -1).T).T
    pw = rays * depth[:, None]                       # in camera-1 = world coordinates
    codes = rng.integers(0, 256, (n_pts, ORB_BYTES), dtype=np.uint8)
    views = []
    for c in range(2):
        Rcw = Rwc[c].T; tcw = -Rcw @ twc[c]
        pc = pw @ Rcw.T + tcw
        pr = pc @ K.T
        u = pr[:, :2] / pr[:, 2:3]
        vis = np.flatnonzero((pc[:, 2] > 0.5) & (u[:, 0] > 10) & (u[:, 0] < img_w - 10) & (u[:, 1] > 10) & (u[:, 1] < img_h - 10))
        sel = vis[rng.permutation(len(vis))[:min(len(vis), n_shared + n_shared // 2)]]
        n_lm = len(sel)
        kp = np.zeros((n_kp, 2), np.float32); octave = rng.integers(0, n_octaves_data, n_kp).astype(np.float32)
        desc = rng.integers(0, 256, (n_kp, ORB_BYTES), dtype=np.uint8)
        kp[:] = np.stack([rng.uniform(0, img_w, n_kp), rng.uniform(0, img_h, n_kp)], -1)
        slots = rng.choice(n_kp, n_lm, replace=False)
        kp[slots] = (u[sel] + rng.normal(0, pix_noise, (n_lm, 2))).astype(np.float32)
        desc[slots] = codes[sel] ^ _flip_mask(rng, (n_lm, ORB_BYTES))
        lm_valid = np.zeros(n_kp, np.uint8); lm_valid[slots] = 1
        lm_valid[slots[rng.random(n_lm) < 0.05]] = 0                        # a few invalid landmarks
        lm_pos = np.zeros((n_kp, 3)); lm_pos[slots] = pw[sel] + rng.normal(0, 0.01, (n_lm, 3))
        lm_maxdist = np.ones(n_kp); lm_maxdist[slots] = np.linalg.norm(pc[sel], axis=1) * rng.uniform(0.9, 2.5, n_lm)
        lm_desc = rng.integers(0, 256, (n_kp, ORB_BYTES), dtype=np.uint8); lm_desc[slots] = codes[sel] ^ _flip_mask(rng, (n_lm, ORB_BYTES))
        Tcw = np.eye(4); Tcw[:3, :3] = Rcw; Tcw[:3, 3] = tcw
        views.append(dict(kp=kp, octave=octave, desc=desc, lm_valid=lm_valid, lm_pos=lm_pos, lm_maxdist=lm_maxdist, lm_desc=lm_desc,
                          K=K.copy(), Tcw=Tcw, img_bounds=np.array([0.0, img_w, 0.0, img_h]), point_of_slot=dict(zip(slots.tolist(), sel.tolist()))))
    # T12 = T_c1_c2 (maps camera-2 coordinates to camera-1 coordinates), perturbed like a RANSAC estimate
    T12 = views[0]["Tcw"] @ np.linalg.inv(views[1]["Tcw"])
    dT = np.eye(4); dT[:3, :3] = _rot(rng.normal(0, pose_noise * 0.2, 3)); dT[:3, 3] = rng.normal(0, pose_noise, 3)
    T12 = T12 @ dT
    T21 = np.linalg.inv(T12)
    already1 = (rng.random(n_kp) < 0.1).astype(np.uint8); already2 = (rng.random(n_kp) < 0.1).astype(np.uint8)
    return views, T12, T21, already1, already2


def relpose_case(seed: int, n: int = 150, outlier_frac: float = 0.1, pix_noise: float = 1.0, cam: dict | None = None):
    """Residual pairs of Optimization::OptimizeRelativePose (optimization_be.cpp:620-831): the same physical points expressed in
    camera A (pA_c) and camera B (pB_c = T12^-1 pA_c + 1 cm noise: the two keyframes own different landmark estimates),
    their observations in both images, octave sigmas, and a perturbed initial T12 = [q, t]."""
    from .synth_map import EUROC_DIST, EUROC_INTR, rot_to_quat
    rng = np.random.default_rng(seed)
    cam = cam or dict(intr=EUROC_INTR, dist=EUROC_DIST, cam_model=0, dist_model=0, xi=0.0)
    R = _rot(rng.normal(0, 0.2, 3)); t = rng.normal(0, 0.3, 3)
    pB = rng.uniform(-3, 3, (n, 3)) + np.array([0, 0, 8.0])
    pA = pB @ R.T + t + rng.normal(0, 0.01, (n, 3))

    def proj(p):
        d = np.asarray(cam["dist"], float)
        den = p[:, 2] + (cam.get("xi", 0.0) * np.linalg.norm(p, axis=1) if cam.get("cam_model", 0) == 1 else 0.0)
        x, y = p[:, 0] / den, p[:, 1] / den
        r2 = x * x + y * y
        if cam.get("dist_model", 0) == 0:
            rad = 1 + d[0] * r2 + d[1] * r2 * r2
            xd = x * rad + 2 * d[2] * x * y + d[3] * (r2 + 2 * x * x); yd = y * rad + d[2] * (r2 + 2 * y * y) + 2 * d[3] * x * y
        elif cam["dist_model"] == 1:
            r = np.sqrt(r2); th = np.arctan(r); t2 = th * th
            s_ = th * (1 + d[0] * t2 + d[1] * t2 ** 2 + d[2] * t2 ** 3 + d[3] * t2 ** 4) / np.maximum(r, 1e-12)
            xd, yd = s_ * x, s_ * y
        else:
            w = d[0]; c = 2 * np.tan(0.5 * w); r = np.sqrt(r2)
            s_ = np.arctan(c * r) / (w * np.maximum(r, 1e-12)); xd, yd = s_ * x, s_ * y
        return np.stack([cam["intr"][0] * xd + cam["intr"][2], cam["intr"][1] * yd + cam["intr"][3]], -1)
    kpA = proj(pA) + rng.normal(0, pix_noise, (n, 2)); kpB = proj(pB) + rng.normal(0, pix_noise, (n, 2))
    out = rng.random(n) < outlier_frac
    kpA[out] += rng.normal(0, 60, (int(out.sum()), 2))
    sA = (rng.integers(0, 3, n) + 1) * 2.0; sB = (rng.integers(0, 3, n) + 1) * 2.0
    T0 = np.concatenate([rot_to_quat(R @ _rot(rng.normal(0, 0.03, 3))), t + rng.normal(0, 0.05, 3)])
    Tgt = np.concatenate([rot_to_quat(R), t])
    return dict(T12=T0, pA_c=pA, pB_c=pB, kpA=kpA.astype(np.float32), kpB=kpB.astype(np.float32), sigmaA=sA, sigmaB=sB, camA=cam, camB=cam), Tgt, out


def projection_search_scene(seed: int, n_kp: int = 1000, n_lm: int = 800, cam: dict | None = None, img_w: int = 752, img_h: int = 480):
    """One keyframe and a list of candidate landmarks ("loop map points") as FeatureMatcher::SearchByProjection sees them
    (feature_matcher_be.cpp:168-291): a third of the landmarks are already observed by the keyframe (at a keypoint with a WORSE
    descriptor than another nearby one for some of them → RemapLandmark), some are invalid / behind the camera / seen from
    behind / out of the distance range, several landmarks compete for the same keypoint."""
    from .synth_map import EUROC_DIST, EUROC_INTR
    rng = np.random.default_rng(seed)
    cam = cam or dict(intr=EUROC_INTR, dist=EUROC_DIST, cam_model=0, dist_model=0, xi=0.0)
    Tcw = np.eye(4); Tcw[:3, :3] = _rot(rng.normal(0, 0.1, 3)); Tcw[:3, 3] = rng.normal(0, 0.2, 3)
    Rwc = Tcw[:3, :3].T; Ow = -Rwc @ Tcw[:3, 3]
    # landmarks in front of the camera (plus a few behind)
    pc = np.stack([rng.uniform(-4, 4, n_lm), rng.uniform(-2.5, 2.5, n_lm), rng.uniform(3, 12, n_lm)], -1)
    pc[rng.random(n_lm) < 0.05, 2] *= -1
    pw = pc @ Rwc.T + Ow                                    # = Rwc pc + Ow
    d = np.asarray(cam["dist"], float)
    den = pc[:, 2] + (cam.get("xi", 0.0) * np.linalg.norm(pc, axis=1) if cam.get("cam_model", 0) == 1 else 0.0)
    x, y = pc[:, 0] / den, pc[:, 1] / den
    r2 = x * x + y * y
    if cam.get("dist_model", 0) == 0:
        rad = 1 + d[0] * r2 + d[1] * r2 * r2
        xd = x * rad + 2 * d[2] * x * y + d[3] * (r2 + 2 * x * x); yd = y * rad + d[2] * (r2 + 2 * y * y) + 2 * d[3] * x * y
    else:
        r = np.sqrt(r2); th = np.arctan(r); t2 = th * th
        s_ = th * (1 + d[0] * t2 + d[1] * t2 ** 2 + d[2] * t2 ** 3 + d[3] * t2 ** 4) / np.maximum(r, 1e-12)
        xd, yd = s_ * x, s_ * y
    uv = np.stack([cam["intr"][0] * xd + cam["intr"][2], cam["intr"][1] * yd + cam["intr"][3]], -1)
    codes = rng.integers(0, 256, (n_lm, ORB_BYTES), dtype=np.uint8)
    kp = np.stack([rng.uniform(0, img_w, n_kp), rng.uniform(0, img_h, n_kp)], -1).astype(np.float32)
    desc = rng.integers(0, 256, (n_kp, ORB_BYTES), dtype=np.uint8)
    octave = np.zeros(n_kp, np.float32)
    inimg = np.flatnonzero((pc[:, 2] > 0) & (uv[:, 0] > 5) & (uv[:, 0] < img_w - 5) & (uv[:, 1] > 5) & (uv[:, 1] < img_h - 5))
    # each visible landmark gets 1-2 keypoints near its projection (a good one and sometimes a decoy with more bit flips); pairs of
    # landmarks share a projection so that they compete for the same keypoint
    slots = rng.permutation(n_kp)
    used = 0
    good_kp = np.full(n_lm, -1)
    for j, l in enumerate(inimg[: n_kp // 3]):
        if j % 7 == 6 and j > 0:
            uv[l] = uv[inimg[j - 1]]; codes[l] = codes[inimg[j - 1]]                 # competitor of the previous landmark
            continue
        k0 = slots[used]; used += 1
        kp[k0] = uv[l] + rng.normal(0, 1.0, 2); desc[k0] = codes[l] ^ _flip_mask(rng, (ORB_BYTES,)); good_kp[l] = k0
        if j % 3 == 0:
            k1 = slots[used]; used += 1
            kp[k1] = uv[l] + rng.normal(0, 3.0, 2); desc[k1] = codes[l] ^ _flip_mask(rng, (ORB_BYTES,)) ^ _flip_mask(rng, (ORB_BYTES,))
    valid = (rng.random(n_lm) > 0.05).astype(np.uint8)
    dist3 = np.linalg.norm(pw - Ow, axis=1)
    normal = (Ow - pw) / dist3[:, None]                     # facing the camera …
    flip = rng.random(n_lm) < 0.05; normal[flip] *= -1      # … except a few seen from behind (viewing-angle gate; PO.Pn >= 0.5 d needed)
    normal = -normal                                        # Pn points from the camera side: PO = p - Ow must align with it
    max_distance = dist3 * rng.uniform(0.9, 1.15, n_lm)
    min_dist = 0.8 * max_distance; max_dist = 1.2 * max_distance
    far = rng.random(n_lm) < 0.05; max_dist[far] = dist3[far] * 0.5            # out of the invariance range
    # the keyframe already observes a third of the visible landmarks: at a random unrelated keypoint (→ remap candidates) or at
    # their good keypoint
    kf_has_lm = np.zeros(n_kp, np.uint8); kf_lm_cand = np.full(n_kp, -1, np.int32); feat_idx = np.full(n_lm, -1, np.int32)
    for j, l in enumerate(inimg[: n_kp // 3: 3]):
        k = good_kp[l] if (j % 2 == 0 and good_kp[l] >= 0) else slots[used + j]
        if kf_has_lm[k]:
            continue
        kf_has_lm[k] = 1; kf_lm_cand[k] = l; feat_idx[l] = k
    extra = slots[-40:]                                     # keypoints holding landmarks that are not in the candidate list
    extra = extra[kf_has_lm[extra] == 0]; kf_has_lm[extra] = 1
    matched = np.zeros(n_kp, np.uint8); matched[rng.choice(n_kp, 30, replace=False)] = 1
    obs_good = [l for l in inimg[: n_kp // 3: 3] if feat_idx[l] >= 0 and feat_idx[l] == good_kp[l]]
    for l in obs_good[::3]:
        matched[feat_idx[l]] = 1                            # existing (better) observation is not a search candidate → bDoNotReplace
    view = dict(kp=kp, octave=octave, desc=desc, lm_valid=kf_has_lm, lm_pos=np.zeros((n_kp, 3)), lm_maxdist=np.ones(n_kp), lm_desc=np.zeros((n_kp, 32), np.uint8),
                K=np.eye(3), Tcw=Tcw, img_bounds=np.array([0.0, img_w, 0.0, img_h]))
    lms = dict(valid=valid, pos=pw, normal=normal, min_dist=min_dist, max_dist=max_dist, max_distance=max_distance, desc=codes ^ _flip_mask(rng, (n_lm, ORB_BYTES)),
               feat_idx=feat_idx)
    return view, kf_lm_cand, Tcw, cam, lms, matched
