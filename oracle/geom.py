"""ctypes binding of oracle/geom_oracle.c (guided search + RANSAC hypothesis scoring) — TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import numpy as np
from . import knn as _k

c_vp = C.c_void_p


class OraKfView(C.Structure):
    _fields_ = [("n", C.c_int32)] + [(k, c_vp) for k in ("kp", "octave", "desc", "lm_valid", "lm_pos", "lm_maxdist", "lm_desc", "grid_ptr", "grid_idx")] + \
               [("grid_w_inv", C.c_double), ("grid_h_inv", C.c_double), ("K", C.c_double * 9), ("Tcw", C.c_double * 16), ("img", C.c_double * 4)]


class OraSearchParams(C.Structure):
    _fields_ = [("th", C.c_double), ("desc_th_low", C.c_int32), ("num_octaves", C.c_int32), ("scale_factor", C.c_double)]


def search_by_se3(kf1, kf2, T12, T21, already1, already2, th=9.5, desc_th_low=50, num_octaves=1, scale_factor=2.0):
    """kf1/kf2: objects with .cstruct(cls) and .n (covins_b200.placerec.KfView builds the flat arrays; the struct layout is
    declared here independently).  → (match12, n_found, match1, match2)"""
    a, b = kf1.cstruct(OraKfView), kf2.cstruct(OraKfView)
    T12 = np.ascontiguousarray(T12, np.float64).reshape(16); T21 = np.ascontiguousarray(T21, np.float64).reshape(16)
    a1 = np.ascontiguousarray(already1, np.uint8); a2 = np.ascontiguousarray(already2, np.uint8)
    prm = OraSearchParams(float(th), int(desc_th_low), int(num_octaves), float(scale_factor))
    m12 = np.full(max(kf1.n, 1), -1, np.int32); m1 = np.full(max(kf1.n, 1), -1, np.int32); m2 = np.full(max(kf2.n, 1), -1, np.int32)
    nf = C.c_int32(0)
    _k.lib().ora_search_by_se3(C.byref(a), C.byref(b), c_vp(T12.ctypes.data), c_vp(T21.ctypes.data), c_vp(a1.ctypes.data), c_vp(a2.ctypes.data), C.byref(prm),
                               c_vp(m12.ctypes.data), C.byref(nf), c_vp(m1.ctypes.data), c_vp(m2.ctypes.data))
    return m12[:kf1.n], int(nf.value), m1[:kf1.n], m2[:kf2.n]


def score_absolute_pose(models, pts, f, sigma, cam_off, cam_rot, threshold):
    m = np.ascontiguousarray(models, np.float64).reshape(-1, 12); p = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    f = np.ascontiguousarray(f, np.float64).reshape(-1, 3); s = np.ascontiguousarray(sigma, np.float64)
    co = np.ascontiguousarray(cam_off, np.float64).reshape(3); cr = np.ascontiguousarray(cam_rot, np.float64).reshape(9)
    H, n = len(m), len(p)
    sc = np.zeros((H, n)); inl = np.zeros((H, n), np.uint8); cnt = np.zeros(max(H, 1), np.int32)
    _k.lib().ora_score_absolute_pose(c_vp(m.ctypes.data), H, c_vp(p.ctypes.data), c_vp(f.ctypes.data), c_vp(s.ctypes.data), n, c_vp(co.ctypes.data),
                                     c_vp(cr.ctypes.data), C.c_double(threshold), c_vp(sc.ctypes.data), c_vp(inl.ctypes.data), c_vp(cnt.ctypes.data))
    return sc, inl, cnt[:H]


def score_relative_pose(models, f1, f2, sigma1, sigma2, threshold):
    m = np.ascontiguousarray(models, np.float64).reshape(-1, 12)
    a = np.ascontiguousarray(f1, np.float64).reshape(-1, 3); b = np.ascontiguousarray(f2, np.float64).reshape(-1, 3)
    s1 = np.ascontiguousarray(sigma1, np.float64); s2 = np.ascontiguousarray(sigma2, np.float64)
    H, n = len(m), len(a)
    sc = np.zeros((H, n)); inl = np.zeros((H, n), np.uint8); cnt = np.zeros(max(H, 1), np.int32)
    _k.lib().ora_score_relative_pose(c_vp(m.ctypes.data), H, c_vp(a.ctypes.data), c_vp(b.ctypes.data), c_vp(s1.ctypes.data), c_vp(s2.ctypes.data), n,
                                     C.c_double(threshold), c_vp(sc.ctypes.data), c_vp(inl.ctypes.data), c_vp(cnt.ctypes.data))
    return sc, inl, cnt[:H]


class OraProjLandmarks(C.Structure):
    _fields_ = [("m", C.c_int32)] + [(k, c_vp) for k in ("valid", "pos", "normal", "min_dist", "max_dist", "max_distance", "desc", "feat_idx")]


def search_by_projection(kf, kf_lm_cand, Tcw, cam, lms, matched, th=10.0, desc_th_low=50, num_octaves=1, scale_factor=2.0):
    """→ (action [m], best_idx [m], n_matches); argument meaning as covins_b200.placerec.search_by_projection"""
    from covins_b200.placerec import _proj_args          # array packing helper only; the struct layouts are declared here
    L, prm, keep, m = _proj_args(kf, kf_lm_cand, Tcw, cam, lms, matched, th, desc_th_low, num_octaves, scale_factor, OraProjLandmarks, OraSearchParams)
    k = kf.cstruct(OraKfView)
    action = np.zeros(max(m, 1), np.int32); best = np.full(max(m, 1), -1, np.int32); nm = C.c_int32(0)
    _k.lib().ora_search_by_projection(C.byref(k), c_vp(keep[1].ctypes.data), c_vp(keep[2].ctypes.data), c_vp(keep[3].ctypes.data), c_vp(keep[4].ctypes.data),
                                      int(cam.get("cam_model", 0)), int(cam.get("dist_model", 0)), C.c_double(float(cam.get("xi", 0.0))), C.byref(L),
                                      c_vp(keep[5].ctypes.data), C.byref(prm), c_vp(action.ctypes.data), c_vp(best.ctypes.data), C.byref(nm))
    return action[:m], best[:m], int(nm.value)
