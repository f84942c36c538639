"""CPU: host logic of the C++ shim that needs no GPU.  covins_b200::SearchByProjection flattens the reference-shaped containers,
makes ONE library call and replays the returned decisions (new match / RemapLandmark) on the containers; here the library call is
answered by a test double backed by the C oracle (tests/cpp/stub_cabi_proj.c), and the resulting container state must equal a
plain Python restatement of the reference's sequential loop (feature_matcher_be.cpp:168-291, keyframe_be.cpp:484-495)."""
import os
import subprocess

import numpy as np
import pytest

from covins_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPP = os.path.join(ROOT, "tests", "cpp")


def _build(tmp):
    exe = os.path.join(tmp, "shim_proj_cpu")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    subprocess.check_call(["gcc", "-O1", "-c", os.path.join(CPP, "stub_cabi_proj.c"), "-o", os.path.join(tmp, "stub.o")])
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-o", exe, os.path.join(CPP, "shim_proj_cpu.cpp"), os.path.join(tmp, "stub.o"),
                           "-L" + os.path.join(ROOT, "oracle"), "-lcovins_oracle", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return exe


def _reference_loop(view, kf_lm_cand, Tcw, cam, lms, matched, th=10.0, th_low=50):
    """the reference's loop with its container side effects → (nmatches, vpMatched codes, landmarks_ codes, feature index per landmark);
    codes: landmark position in vpPoints, -1 = nullptr, -2 = a landmark outside the list"""
    from test_geom import _py_grid
    grid = _py_grid(view["kp"])
    n, m = len(view["kp"]), len(lms["pos"])
    vp_matched = np.where(matched > 0, -2, -1).astype(np.int32)
    slot = np.where(view["lm_valid"] > 0, np.where(kf_lm_cand >= 0, kf_lm_cand, -2), -1).astype(np.int32)     # landmarks_
    feat = lms["feat_idx"].copy()
    bits = np.unpackbits(view["desc"], axis=1)
    R, t = Tcw[:3, :3], Tcw[:3, 3]; Ow = -R.T @ t
    d, intr = np.asarray(cam["dist"]), np.asarray(cam["intr"])
    nm = 0
    for i in range(m):
        if not lms["valid"][i]:
            continue
        pw = lms["pos"][i]; pc = R @ pw + t
        if pc[2] < 0:
            continue
        x, y = pc[0] / pc[2], pc[1] / pc[2]; r2 = x * x + y * y; rad = 1 + d[0] * r2 + d[1] * r2 * r2
        u = intr[0] * (x * rad + 2 * d[2] * x * y + d[3] * (r2 + 2 * x * x)) + intr[2]
        v = intr[1] * (y * rad + d[2] * (r2 + 2 * y * y) + 2 * d[3] * x * y) + intr[3]
        if not (0 <= u < 752 and 0 <= v < 480):
            continue
        PO = pw - Ow; d3 = np.linalg.norm(PO)
        if d3 < lms["min_dist"][i] or d3 > lms["max_dist"][i] or PO @ lms["normal"][i] < 0.5 * d3:
            continue
        tx, ty = np.float32(u), np.float32(v)
        cx0 = max(0, int(np.floor((float(tx) - th) * 64 / 752))); cx1 = min(63, int(np.ceil((float(tx) + th) * 64 / 752)))
        cy0 = max(0, int(np.floor((float(ty) - th) * 48 / 480))); cy1 = min(47, int(np.ceil((float(ty) + th) * 48 / 480)))
        lb = np.unpackbits(lms["desc"][i]); ham = lambda k: int((bits[k] != lb).sum())
        bd, best = 256, -1
        for ix in range(cx0, cx1 + 1):
            for iy in range(cy0, cy1 + 1):
                for idx in grid.get((ix, iy), []):
                    dx = np.float32(view["kp"][idx, 0]) - tx; dy = np.float32(view["kp"][idx, 1]) - ty
                    if float(np.sqrt(np.float32(dx * dx + dy * dy))) > th or vp_matched[idx] != -1:
                        continue
                    dd = ham(idx)
                    if dd < bd:
                        bd, best = dd, idx
        if best < 0 or bd > th_low:
            continue
        ex = feat[i]
        if ex != -1:
            if ham(ex) < bd or (slot[best] != -1 and ham(best) < bd):
                continue
            displaced = slot[best]                                   # RemapLandmark (keyframe_be.cpp:484-495)
            slot[ex] = -1; slot[best] = i; feat[i] = best
            if displaced >= 0:
                feat[displaced] = -1
        else:
            vp_matched[best] = i; nm += 1
    return nm, vp_matched, slot, feat


@pytest.mark.parametrize("seed", [0, 4])
def test_shim_search_by_projection_replays_the_reference_loop(tmp_path, seed):
    exe = _build(str(tmp_path))
    view, kf_lm_cand, Tcw, cam, lms, matched = synth.projection_search_scene(seed, n_kp=600, n_lm=500)
    # the mock landmark derives its invariance range from min/max_distance_ as the reference does (x0.8 / x1.2)
    lms["max_distance"] = np.where(lms["max_dist"] < lms["max_distance"], lms["max_dist"] / 1.2, lms["max_distance"])
    lms["max_dist"] = 1.2 * lms["max_distance"]
    min_distance = lms["min_dist"] / 0.8
    lms["min_dist"] = 0.8 * min_distance
    d = str(tmp_path)
    for name, a, dt in (("kf_kp", view["kp"], np.float32), ("kf_octave", view["octave"], np.float32), ("kf_desc", view["desc"], np.uint8),
                        ("kf_has_lm", view["lm_valid"], np.uint8), ("matched", matched, np.uint8), ("kf_lm_cand", kf_lm_cand, np.int32),
                        ("lm_feat_idx", lms["feat_idx"], np.int32), ("Tcw", Tcw, np.float64), ("intr", cam["intr"], np.float64),
                        ("dist", cam["dist"], np.float64), ("lm_valid", lms["valid"], np.uint8), ("lm_desc", lms["desc"], np.uint8),
                        ("lm_pos", lms["pos"], np.float64), ("lm_normal", lms["normal"], np.float64), ("lm_min_distance", min_distance, np.float64),
                        ("lm_max_distance", lms["max_distance"], np.float64)):
        np.ascontiguousarray(a, dt).tofile(os.path.join(d, name + ".bin"))
    subprocess.check_call([exe, d])
    out = np.fromfile(os.path.join(d, "proj_out.bin"), np.int32)
    n, m = len(view["kp"]), len(lms["pos"])
    nm, vp_matched, slot, feat = _reference_loop(view, kf_lm_cand, Tcw, cam, lms, matched)
    assert out[0] == nm and nm > 30
    assert np.array_equal(out[1:1 + n], vp_matched)
    assert np.array_equal(out[1 + n:1 + 2 * n], slot)
    assert np.array_equal(out[1 + 2 * n:1 + 2 * n + m], feat)
    assert (slot != np.where(view["lm_valid"] > 0, np.where(kf_lm_cand >= 0, kf_lm_cand, -2), -1)).sum() >= 4      # remaps happened
