"""ctypes binding of oracle/ba_port.cpp — the compiled, threaded CPU port of the optimisation hot path (TEST
INFRASTRUCTURE and the timed CPU baseline of bench.py; never imported by covins_b200/).  Same flat problem dict as
oracle/ba_oracle.py / covins_b200.synth_map."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
c_vp = C.c_void_p

TERMINATION = {0: "NO_CONVERGENCE", 1: "CONVERGENCE(gradient)", 2: "CONVERGENCE(parameter)", 3: "CONVERGENCE(function)", 4: "FAILURE"}
STEP = {1: "accepted", 2: "rejected", 3: "invalid", 4: "converged"}
GBA_LOOP_SQRT_INFO = np.diag([100.0] * 3 + [1e4] * 3)   # optimization_be.cpp:238-240, 534-536


class Problem(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("K", "L", "n_obs", "n_imu", "n_edge", "n_cam")] + [(n, c_vp) for n in (
        "pose", "speedbias", "pose_const", "cam_of_kf", "extr", "intr", "dist", "lm", "lm_obs_ptr", "obs_kf", "obs_uv", "obs_sigma",
        "obs_skip", "imu_i", "imu_j", "imu_ptr", "imu_dt", "imu_acc", "imu_gyr", "imu_acc0", "imu_gyr0", "imu_noise", "edge_i",
        "edge_j", "edge_q", "edge_t", "edge_sqrt_info", "edge_robust")]


class Options(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("visual_only", C.c_int32), ("cauchy_reproj", C.c_double), ("cauchy_edge", C.c_double),
                ("threads", C.c_int32)]


class Result(C.Structure):
    _fields_ = [("pose", c_vp), ("speedbias", c_vp), ("lm", c_vp), ("cost_history", c_vp), ("step_status", c_vp),
                ("cost_history_cap", C.c_int32), ("n_cost_history", C.c_int32), ("iterations", C.c_int32), ("termination", C.c_int32),
                ("initial_cost", C.c_double), ("final_cost", C.c_double), ("phase_s", C.c_double * 6), ("factor_flops", C.c_double)]


def blas_path():
    import scipy
    c = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so")))
    return c[0] if c else None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libcovins_ba_port.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        _LIB = C.CDLL(path)
        bp = blas_path()
        _LIB.blas = bool(bp) and _LIB.bap_init_blas(bp.encode()) == 0
    return _LIB


def _arr(a, dt):
    return None if a is None else np.ascontiguousarray(a, dtype=dt)


class _Flat:
    def __init__(self, p, edges=None, obs_skip=None, use_imu=True, loop_robust=True):
        K = int(p["K"]); L = int(p.get("L", 0))
        self.a = a = {}
        a["pose"] = _arr(p["pose"], np.float64); a["speedbias"] = _arr(p.get("speedbias"), np.float64)
        a["pose_const"] = _arr(p["pose_const"], np.uint8); a["cam_of_kf"] = _arr(p.get("cam_of_kf"), np.int32)
        a["extr"] = _arr(p["extr"], np.float64).reshape(-1, 7)
        a["intr"] = _arr(p.get("intr", np.zeros((len(a["extr"]), 4))), np.float64)
        a["dist"] = _arr(p.get("dist", np.zeros((len(a["extr"]), 4))), np.float64)
        a["lm"] = _arr(p.get("lm", np.zeros((0, 3))), np.float64); a["lm_obs_ptr"] = _arr(p.get("lm_obs_ptr", np.zeros(1)), np.int32)
        a["obs_kf"] = _arr(p.get("obs_kf", np.zeros(0)), np.int32); a["obs_uv"] = _arr(p.get("obs_uv", np.zeros((0, 2))), np.float32)
        a["obs_sigma"] = _arr(p.get("obs_sigma", np.zeros(0)), np.float64); a["obs_skip"] = _arr(obs_skip, np.uint8)
        n_imu = len(p.get("imu_i", [])) if use_imu else 0
        for k, dt in (("imu_i", np.int32), ("imu_j", np.int32), ("imu_ptr", np.int32), ("imu_dt", np.float64), ("imu_acc", np.float64),
                      ("imu_gyr", np.float64), ("imu_acc0", np.float64), ("imu_gyr0", np.float64), ("imu_noise", np.float64)):
            a[k] = _arr(p.get(k), dt) if n_imu else None
        if edges is None and len(p.get("loop_i", [])) > 0:
            n = len(p["loop_i"])
            edges = dict(i=p["loop_i"], j=p["loop_j"], q=p["loop_q"], t=p["loop_t"], sqrt_info=np.tile(GBA_LOOP_SQRT_INFO[None], (n, 1, 1)),
                         robust=np.full(n, 1 if loop_robust else 0, np.uint8))
        n_edge = 0
        if edges is not None and len(edges["i"]) > 0:
            n_edge = len(edges["i"])
            a["edge_i"] = _arr(edges["i"], np.int32); a["edge_j"] = _arr(edges["j"], np.int32)
            a["edge_q"] = _arr(edges["q"], np.float64); a["edge_t"] = _arr(edges["t"], np.float64)
            a["edge_sqrt_info"] = _arr(np.asarray(edges["sqrt_info"]).reshape(n_edge, 36), np.float64)
            a["edge_robust"] = _arr(edges["robust"], np.uint8)
        s = Problem()
        s.K, s.L, s.n_obs, s.n_imu, s.n_edge, s.n_cam = K, L, len(a["obs_kf"]), n_imu, n_edge, len(a["extr"])
        for name, _ in Problem._fields_[6:]:
            v = a.get(name)
            setattr(s, name, v.ctypes.data if v is not None else None)
        self.s, self.K, self.L, self.n_obs = s, K, L, s.n_obs


class _Res:
    def __init__(self, K, L, cap=64):
        self.pose = np.zeros((K, 7)); self.sb = np.zeros((K, 9)); self.lm = np.zeros((max(L, 1), 3))
        self.hist = np.zeros(cap); self.status = np.zeros(cap, np.uint8)
        r = Result()
        r.pose, r.speedbias, r.lm = self.pose.ctypes.data, self.sb.ctypes.data, self.lm.ctypes.data
        r.cost_history, r.step_status, r.cost_history_cap = self.hist.ctypes.data, self.status.ctypes.data, cap
        self.r, self.L = r, L

    def as_dict(self):
        r = self.r; n = r.n_cost_history
        ph = list(r.phase_s)
        return dict(pose=self.pose, speedbias=self.sb, lm=self.lm[:self.L], cost=self.hist[:n].copy(),
                    steps=[STEP.get(int(x), "?") for x in self.status[:max(n - 1, 0)]], iterations=int(r.iterations),
                    termination=TERMINATION.get(int(r.termination), "?"), initial_cost=float(r.initial_cost), final_cost=float(r.final_cost),
                    phase_s=dict(linearize=ph[0], build_schur=ph[1], factor=ph[2], solve=ph[3], step=ph[4], setup=ph[5]),
                    factor_flops=float(r.factor_flops))


def solve(p, max_iterations, visual_only=False, cauchy_reproj=1.0, cauchy_edge=1.0, edges=None, obs_skip=None, threads=0,
          loop_robust=True):
    """one ceres::Solve equivalent (GBA round / PGO) on the CPU"""
    f = _Flat(p, edges=edges, obs_skip=obs_skip, use_imu=not visual_only, loop_robust=loop_robust)
    o = Options(int(max_iterations), int(visual_only), float(cauchy_reproj), float(cauchy_edge), int(threads))
    res = _Res(f.K, f.L)
    rc = lib().bap_solve(C.byref(f.s), C.byref(o), C.byref(res.r))
    if rc:
        raise RuntimeError(f"bap_solve failed ({rc})")
    return res.as_dict()


def global_bundle_adjustment(p, iterations_limit=10, visual_only=False, outlier_removal=True, th_outlier=0.92, threads=0):
    f = _Flat(p, use_imu=not visual_only)
    res = _Res(f.K, f.L)
    removed = np.zeros(max(f.n_obs, 1), np.uint8)
    rc = lib().bap_gba(C.byref(f.s), int(iterations_limit), int(visual_only), int(outlier_removal), C.c_double(th_outlier), int(threads),
                       C.byref(res.r), C.c_void_p(removed.ctypes.data))
    if rc:
        raise RuntimeError(f"bap_gba failed ({rc})")
    d = res.as_dict(); d["obs_removed"] = removed[:f.n_obs].astype(bool)
    return d
