"""oracle/ba_oracle.py — CPU restatement of the reference's optimisation half (PGO / global BA).

TEST INFRASTRUCTURE ONLY (imported from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg).

PARITY UNPINNED.  The reference builds a ceres::Problem in
covins_backend/src/covins_backend/optimization_be.cpp (GlobalBundleAdjustment :56-618, PoseGraphOptimization
:833-1086) and hands all arithmetic to third-party code that is NOT in /root/reference and cannot be built
offline: robopt_open @ branch fix_imu_residual (no commit pin, dependencies.rosinstall:67-70), ceres_catkin
(no pin, :39-41), aslam_cv2 (:63-65).  The reference has no tests, fixtures or golden vectors for this path
(SURVEY.md §4, §8c).  What follows restates the published algorithms of those libraries; every convention
marked [A] is an assumption.  The CUDA path is checked against THIS file, and this
file is checked by finite differences / invariants (tests/test_oracle_ba.py) — not against Ceres.

Deliberately independent of the CUDA implementation: Jacobians here come from torch autograd (fp64) of the
plain residual formulas, the normal equations from scipy.sparse, the factorisation from LAPACK.

Restated pieces
  problem construction   optimization_be.cpp:296-557 (GBA round 2), :62-254 (round 1), :833-1031 (PGO)
  state layout           keyframe_base.cpp:486-521: pose [qx,qy,qz,qw,x,y,z] = T_ws, speed-bias [v,ba,bg]
  Plus [A]               PoseQuaternionLocalParameterization: delta = [dtheta, dp], q+ = Exp(dtheta) * q, p+ = p + dp
  reprojection [A]       GlobalEuclideanReprError<Pinhole,RadTan>: r = (pi(T_sc^-1 T_ws^-1 p_w) - kp) / sigma; a point
                         behind the camera (z < 1e-10, aslam POINT_BEHIND_CAMERA) gives zero residual and Jacobian;
                         sigma = 2 (octave + 1) (optimization_be.cpp:183-184, 477-478)
  between [A]            SixDofBetweenError(kImu): e = sqrt_info [2 vec(q_m^-1 q_1^-1 q_2); R_1^T (t_2 - t_1) - t_m],
                         rotation first (pinned by how :896-897 / :239-240 fill sqrt_info)
  IMU [A]                PreintegrationBase = VINS-Mono IntegrationBase (midpoint), residual
                         [dp, dtheta, dv, dba, dbg] whitened by chol(P^-1)^T; repropagate at the current bias
                         (optimization_be.cpp:132-140, 387-396); bias correction uses the exact Exp map
  loss                   ceres::CauchyLoss(a) + Corrector (rho'' <= 0 → scale r and J by sqrt(rho')); cost = rho/2
  solver                 Ceres 1.x TrustRegionMinimizer + DoglegStrategy(TRADITIONAL_DOGLEG), Jacobi scaling,
                         defaults of Ceres 1.14 [A]; linear solve = exact solution of (J^T J + D^2) x = J^T r
                         (SPARSE_SCHUR computes the same x; landmarks are eliminated first here too)
"""
from __future__ import annotations

import math

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import torch

torch.set_default_dtype(torch.float64)
F64 = torch.float64


# ------------------------------------------------------------------------------------------------ quaternions (x,y,z,w)
def qmul(a, b):
    ax, ay, az, aw = a.unbind(-1)
    bx, by, bz, bw = b.unbind(-1)
    return torch.stack([aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw,
                        aw * bw - ax * bx - ay * by - az * bz], -1)


def qconj(q):
    return torch.cat([-q[..., :3], q[..., 3:]], -1)


def qrot(q, v):
    """R(q) v"""
    u, w = q[..., :3], q[..., 3:]
    t = 2 * torch.cross(u, v, dim=-1)
    return v + w * t + torch.cross(u, t, dim=-1)


def qexp(phi):
    """Exp: rotation vector → unit quaternion, smooth at 0 (autograd safe)."""
    th2 = (phi * phi).sum(-1, keepdim=True)
    small = th2 < 1e-12
    th = torch.sqrt(torch.where(small, torch.ones_like(th2), th2))
    k = torch.where(small, 0.5 - th2 / 48.0, torch.sin(0.5 * th) / th)
    w = torch.where(small, 1.0 - th2 / 8.0, torch.cos(0.5 * th))
    return torch.cat([k * phi, w], -1)


def pose_plus(pose, delta):
    q = qmul(qexp(delta[..., :3]), pose[..., :4])
    q = q / q.norm(dim=-1, keepdim=True)
    return torch.cat([q, pose[..., 4:] + delta[..., 3:]], -1)


# ------------------------------------------------------------------------------------------------ residuals
def reproj_residual(pose, lm, extr, intr, dist, uv, sigma, cam_model=None, dist_model=None, xi=None):
    """GlobalEuclideanReprError<Camera, Distortion> [A]: Camera 0 pinhole / 1 unified (xi), Distortion 0 radtan / 1 equidistant
    / 2 fisheye (FOV) — per observation (tensors) or None = pinhole + radtan (optimization_be.cpp:186-231)."""
    q_ws, t_ws = pose[..., :4], pose[..., 4:]
    p_s = qrot(qconj(q_ws), lm - t_ws)
    p_c = qrot(qconj(extr[..., :4]), p_s - extr[..., 4:])
    n = p_c.shape[0]
    cam = torch.zeros(n, dtype=torch.long) if cam_model is None else cam_model
    dm = torch.zeros(n, dtype=torch.long) if dist_model is None else dist_model
    xi = torch.zeros(n) if xi is None else xi
    X, Y, Z = p_c.unbind(-1)
    d = p_c.norm(dim=-1)
    den = torch.where(cam == 1, Z + xi * d, Z)
    # aslam::ProjectionResult POINT_BEHIND_CAMERA (z < 1e-10) / outside the unified model: the error term zeroes residual
    # and Jacobians [A]
    front = den > 1e-10
    dens = torch.where(front, den, torch.ones_like(den))
    x, y = X / dens, Y / dens
    r2 = x * x + y * y
    k1, k2, p1, p2 = dist.unbind(-1)
    # radtan
    rad = 1 + k1 * r2 + k2 * r2 * r2
    xd0 = x * rad + 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
    yd0 = y * rad + p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
    # equidistant (k1..k4)
    rs = torch.sqrt(torch.where(r2 > 1e-16, r2, torch.ones_like(r2)))
    th = torch.atan(rs); t2 = th * th
    thd = th * (1 + k1 * t2 + k2 * t2 * t2 + p1 * t2 ** 3 + p2 * t2 ** 4)
    s1 = torch.where(r2 > 1e-16, thd / rs, torch.ones_like(rs))
    # fisheye / FOV (w = k1)
    w = k1
    ws = torch.where(w * w < 1e-5, torch.ones_like(w), w)
    c = 2 * torch.tan(0.5 * ws)
    rs2 = torch.sqrt(torch.where(r2 >= 1e-5, r2, torch.ones_like(r2)))
    s2 = torch.where(w * w < 1e-5, torch.ones_like(w), torch.where(r2 < 1e-5, c / ws, torch.atan(c * rs2) / (ws * rs2)))
    xd = torch.where(dm == 0, xd0, torch.where(dm == 1, s1 * x, s2 * x))
    yd = torch.where(dm == 0, yd0, torch.where(dm == 1, s1 * y, s2 * y))
    u = intr[..., 0] * xd + intr[..., 2]
    v = intr[..., 1] * yd + intr[..., 3]
    r = torch.stack([(u - uv[..., 0]) / sigma, (v - uv[..., 1]) / sigma], -1)
    return torch.where(front[..., None], r, torch.zeros_like(r))


def between_residual(pose1, pose2, q_m, t_m, sqrt_info):
    q1, t1, q2, t2 = pose1[..., :4], pose1[..., 4:], pose2[..., :4], pose2[..., 4:]
    q12 = qmul(qconj(q1), q2)
    e_rot = 2 * qmul(qconj(q_m), q12)[..., :3]
    e_t = qrot(qconj(q1), t2 - t1) - t_m
    e = torch.cat([e_rot, e_t], -1)
    return torch.einsum("nij,nj->ni", sqrt_info, e)


def imu_residual(pose_i, sb_i, pose_j, sb_j, pre):
    """pre: dict of tensors per factor: dt_sum, alpha, beta, gamma(quat), J (15x15), sqrt_info (15x15), ba_lin, bg_lin, g."""
    qi, pi, qj, pj = pose_i[..., :4], pose_i[..., 4:], pose_j[..., :4], pose_j[..., 4:]
    vi, bai, bgi = sb_i[..., :3], sb_i[..., 3:6], sb_i[..., 6:]
    vj, baj, bgj = sb_j[..., :3], sb_j[..., 3:6], sb_j[..., 6:]
    T = pre["dt_sum"][..., None]
    gvec = torch.zeros_like(pi); gvec[..., 2] = pre["g"]
    dba, dbg = bai - pre["ba_lin"], bgi - pre["bg_lin"]
    J = pre["J"]
    a_hat = pre["alpha"] + torch.einsum("nij,nj->ni", J[:, 0:3, 9:12], dba) + torch.einsum("nij,nj->ni", J[:, 0:3, 12:15], dbg)
    b_hat = pre["beta"] + torch.einsum("nij,nj->ni", J[:, 6:9, 9:12], dba) + torch.einsum("nij,nj->ni", J[:, 6:9, 12:15], dbg)
    g_hat = qmul(pre["gamma"], qexp(torch.einsum("nij,nj->ni", J[:, 3:6, 12:15], dbg)))
    r_p = qrot(qconj(qi), 0.5 * gvec * T * T + pj - pi - vi * T) - a_hat
    r_q = 2 * qmul(qconj(g_hat), qmul(qconj(qi), qj))[..., :3]
    r_v = qrot(qconj(qi), gvec * T + vj - vi) - b_hat
    r = torch.cat([r_p, r_q, r_v, baj - bai, bgj - bgi], -1)
    return torch.einsum("nij,nj->ni", pre["sqrt_info"], r)


# ------------------------------------------------------------------------------------------------ preintegration
def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _q2R(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _qmul_np(a, b):
    ax, ay, az, aw = a; bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw, aw * bw - ax * bx - ay * by - az * bz])


def repropagate(dt, acc, gyr, acc0, gyr0, ba, bg, noise):
    """VINS-Mono IntegrationBase::repropagate / midPointIntegration [A]: returns (dt_sum, alpha, beta, gamma, J, P)."""
    sa, sg, saw, sgw = noise[:4]
    Q = np.diag(np.repeat([sa * sa, sg * sg, sa * sa, sg * sg, saw * saw, sgw * sgw], 3))
    dp, dv, dq = np.zeros(3), np.zeros(3), np.array([0, 0, 0, 1.0])
    Jm, P = np.eye(15), np.zeros((15, 15))
    a0, g0, T = acc0.copy(), gyr0.copy(), 0.0
    I3 = np.eye(3)
    for k in range(len(dt)):
        h, a1, g1 = dt[k], acc[k], gyr[k]
        R0 = _q2R(dq)
        un_acc_0 = R0 @ (a0 - ba)
        un_gyr = 0.5 * (g0 + g1) - bg
        q1 = _qmul_np(dq, np.array([un_gyr[0] * h / 2, un_gyr[1] * h / 2, un_gyr[2] * h / 2, 1.0]))
        q1 = q1 / np.linalg.norm(q1)
        R1 = _q2R(q1)
        un_acc_1 = R1 @ (a1 - ba)
        un_acc = 0.5 * (un_acc_0 + un_acc_1)
        ndp = dp + dv * h + 0.5 * un_acc * h * h
        ndv = dv + un_acc * h
        Rw, Ra0, Ra1 = _skew(un_gyr), _skew(a0 - ba), _skew(a1 - ba)
        F = np.zeros((15, 15)); V = np.zeros((15, 18))
        F[0:3, 0:3] = I3
        F[0:3, 3:6] = -0.25 * R0 @ Ra0 * h * h - 0.25 * R1 @ Ra1 @ (I3 - Rw * h) * h * h
        F[0:3, 6:9] = I3 * h
        F[0:3, 9:12] = -0.25 * (R0 + R1) * h * h
        F[0:3, 12:15] = -0.25 * R1 @ Ra1 * h * h * (-h)
        F[3:6, 3:6] = I3 - Rw * h
        F[3:6, 12:15] = -I3 * h
        F[6:9, 3:6] = -0.5 * R0 @ Ra0 * h - 0.5 * R1 @ Ra1 @ (I3 - Rw * h) * h
        F[6:9, 6:9] = I3
        F[6:9, 9:12] = -0.5 * (R0 + R1) * h
        F[6:9, 12:15] = -0.5 * R1 @ Ra1 * h * (-h)
        F[9:12, 9:12] = I3; F[12:15, 12:15] = I3
        V[0:3, 0:3] = 0.25 * R0 * h * h
        V[0:3, 3:6] = 0.25 * (-R1 @ Ra1 * h * h) * 0.5 * h
        V[0:3, 6:9] = 0.25 * R1 * h * h
        V[0:3, 9:12] = V[0:3, 3:6]
        V[3:6, 3:6] = 0.5 * I3 * h; V[3:6, 9:12] = 0.5 * I3 * h
        V[6:9, 0:3] = 0.5 * R0 * h
        V[6:9, 3:6] = 0.5 * (-R1 @ Ra1 * h) * 0.5 * h
        V[6:9, 6:9] = 0.5 * R1 * h
        V[6:9, 9:12] = V[6:9, 3:6]
        V[9:12, 12:15] = I3 * h; V[12:15, 15:18] = I3 * h
        Jm = F @ Jm
        P = F @ P @ F.T + V @ Q @ V.T
        dp, dv, dq = ndp, ndv, q1
        a0, g0, T = a1, g1, T + h
    return T, dp, dv, dq, Jm, P


# ------------------------------------------------------------------------------------------------ problem
class Problem:
    """Flat problem (covins_b200.synth_map format) → residual blocks with autograd Jacobians.

    options: visual_only (opt.cpp:90,332), use_imu, loop edges with/without loss (round 1 no loss :253,
    round 2 Cauchy(1) :555), reprojection Cauchy(1.0) (:68,302).
    `edges` (PGO): dict(i, j, q, t, sqrt_info, robust) generic between edges (opt.cpp:895-1021).
    """

    def __init__(self, p, visual_only=False, cauchy_reproj=1.0, loop_loss=None, loop_sqrt_info=None, edges=None,
                 use_obs=None, cauchy_edges=0.5):
        self.K, self.L = int(p["K"]), int(p["L"]) if "L" in p else 0
        self.visual_only = visual_only
        self.pose = torch.tensor(np.asarray(p["pose"], float))
        self.sb = torch.tensor(np.asarray(p.get("speedbias", np.zeros((self.K, 9))), float))
        self.lm = torch.tensor(np.asarray(p.get("lm", np.zeros((0, 3))), float)).reshape(-1, 3)
        self.const = np.asarray(p["pose_const"]).astype(bool)
        cam = np.asarray(p.get("cam_of_kf", np.zeros(self.K, np.int32)))
        self.extr_kf = torch.tensor(np.asarray(p["extr"], float))[cam]
        self.cauchy_reproj = cauchy_reproj
        # ---- observations (CSR by landmark) ----
        if self.L > 0:
            ptr = np.asarray(p["lm_obs_ptr"])
            self.obs_lm = np.repeat(np.arange(self.L), np.diff(ptr))
            self.obs_kf = np.asarray(p["obs_kf"]).astype(np.int64)
            keep = np.ones(len(self.obs_kf), bool) if use_obs is None else np.asarray(use_obs).astype(bool)
            self.obs_keep = keep
            # landmarks with < 2 observations are not in the problem (opt.cpp:158-171, 438-453)
            cnt = np.bincount(self.obs_lm[keep], minlength=self.L)
            self.lm_in = cnt >= 2
            sel = keep & self.lm_in[self.obs_lm]
            self.obs_sel = np.flatnonzero(sel)
            self.o_lm = self.obs_lm[sel]; self.o_kf = self.obs_kf[sel]
            self.o_uv = torch.tensor(np.asarray(p["obs_uv"], float)[sel])
            self.o_sigma = torch.tensor(np.asarray(p["obs_sigma"], float)[sel])
            self.o_intr = torch.tensor(np.asarray(p["intr"], float))[cam[self.o_kf]]
            self.o_dist = torch.tensor(np.asarray(p["dist"], float))[cam[self.o_kf]]
            ncam = len(np.asarray(p["extr"]).reshape(-1, 7))
            self.o_cam = torch.tensor(np.asarray(p["cam_model"] if p.get("cam_model") is not None else np.zeros(ncam), np.int64))[cam[self.o_kf]]
            self.o_dm = torch.tensor(np.asarray(p["dist_model"] if p.get("dist_model") is not None else np.zeros(ncam), np.int64))[cam[self.o_kf]]
            self.o_xi = torch.tensor(np.asarray(p["cam_xi"] if p.get("cam_xi") is not None else np.zeros(ncam), float))[cam[self.o_kf]]
        else:
            self.lm_in = np.zeros(0, bool); self.o_lm = np.zeros(0, np.int64); self.o_kf = np.zeros(0, np.int64)
        # ---- IMU factors ----
        self.imu = None
        if not visual_only and len(p.get("imu_i", [])) > 0:
            n = len(p["imu_i"]); ptr = p["imu_ptr"]
            pre = {k: [] for k in ("dt_sum", "alpha", "beta", "gamma", "J", "sqrt_info", "ba_lin", "bg_lin")}
            for f in range(n):
                j = int(p["imu_j"][f])
                ba, bg = np.asarray(p["speedbias"][j][3:6], float), np.asarray(p["speedbias"][j][6:9], float)
                s, e = ptr[f], ptr[f + 1]
                if e - s == 0:
                    raise ValueError("IMU factor with 0 measurements must be dropped by the caller (opt.cpp:382-385)")
                T, a, b, g, Jm, P = repropagate(p["imu_dt"][s:e], p["imu_acc"][s:e], p["imu_gyr"][s:e], p["imu_acc0"][f],
                                                p["imu_gyr0"][f], ba, bg, p["imu_noise"])
                Lc = np.linalg.cholesky(np.linalg.inv(P))
                for k, v in zip(pre, (T, a, b, g, Jm, Lc.T, ba, bg)):
                    pre[k].append(v)
            self.imu = {k: torch.tensor(np.array(v)) for k, v in pre.items()}
            self.imu["g"] = float(p["imu_noise"][4])
            self.imu_i = np.asarray(p["imu_i"]).astype(np.int64); self.imu_j = np.asarray(p["imu_j"]).astype(np.int64)
        # ---- between edges: GBA loop edges or PGO edges ----
        self.edges = None
        if edges is not None:
            self.edges = dict(i=np.asarray(edges["i"]).astype(np.int64), j=np.asarray(edges["j"]).astype(np.int64),
                              q=torch.tensor(np.asarray(edges["q"], float)), t=torch.tensor(np.asarray(edges["t"], float)),
                              S=torch.tensor(np.asarray(edges["sqrt_info"], float).reshape(-1, 6, 6)),
                              robust=np.asarray(edges["robust"]).astype(bool), a=cauchy_edges)
        elif len(p.get("loop_i", [])) > 0:
            n = len(p["loop_i"])
            S = np.tile(np.diag([100.0] * 3 + [1e4] * 3)[None], (n, 1, 1)) if loop_sqrt_info is None else loop_sqrt_info
            self.edges = dict(i=np.asarray(p["loop_i"]).astype(np.int64), j=np.asarray(p["loop_j"]).astype(np.int64),
                              q=torch.tensor(np.asarray(p["loop_q"], float)), t=torch.tensor(np.asarray(p["loop_t"], float)),
                              S=torch.tensor(S), robust=np.full(n, loop_loss is not None), a=loop_loss or 1.0)
        # ---- local parameter layout: [kf0: pose6 (+sb9)] ... then landmarks (3 each) ----
        per = 6 if visual_only else 15
        self.per = per
        self.col_pose = np.where(self.const, -1, np.arange(self.K) * per)
        self.col_sb = np.full(self.K, -1) if visual_only else np.arange(self.K) * per + 6
        self.ncam = self.K * per
        lm_ids = np.flatnonzero(self.lm_in)
        self.col_lm = np.full(self.L, -1); self.col_lm[lm_ids] = self.ncam + 3 * np.arange(len(lm_ids))
        self.n = self.ncam + 3 * len(lm_ids)
        # columns of constant poses stay in the index space but are never written (zero columns → removed below)
        self.active = np.ones(self.n, bool)
        for k in np.flatnonzero(self.const):
            self.active[k * per:k * per + 6] = False

    # -- evaluate all residual blocks at (pose, sb, lm); with_jac → scipy CSR of the corrected Jacobian
    def evaluate(self, pose, sb, lm, with_jac=True):
        rows, cols, vals, res, cost = [], [], [], [], 0.0
        row0 = 0
        info = {}

        def add_block(rfun, params, colidx, dims, loss_a):
            """params: list of tensors [n, *]; colidx: list of int arrays (start col or -1) ; dims local dims"""
            nonlocal row0, cost
            n = params[0].shape[0]
            deltas = [torch.zeros(n, d, requires_grad=with_jac) for d in dims]
            r = rfun(deltas)
            m = r.shape[1]
            s = (r.detach() ** 2).sum(1)
            if loss_a is None:
                scale = torch.ones(n); c = 0.5 * s
            else:
                a2 = loss_a if torch.is_tensor(loss_a) else torch.full((n,), float(loss_a) ** 2)
                robust = torch.isfinite(a2)
                a2s = torch.where(robust, a2, torch.ones_like(a2))
                rho1 = torch.where(robust, 1.0 / (1.0 + s / a2s), torch.ones_like(s))
                c = torch.where(robust, 0.5 * a2s * torch.log1p(s / a2s), 0.5 * s)
                scale = torch.sqrt(rho1)
            cost += float(c.sum())
            res.append((r.detach() * scale[:, None]).reshape(-1))
            if with_jac:
                for comp in range(m):
                    grads = torch.autograd.grad(r[:, comp].sum(), deltas, retain_graph=True, allow_unused=True)
                    for g, ci, d in zip(grads, colidx, dims):
                        if g is None:
                            continue
                        ok = ci >= 0
                        if not ok.any():
                            continue
                        gv = (g * scale[:, None]).numpy()[ok]
                        rr = (row0 + np.arange(n) * m + comp)[ok]
                        rows.append(np.repeat(rr, d)); cols.append((ci[ok][:, None] + np.arange(d)[None, :]).reshape(-1))
                        vals.append(gv.reshape(-1))
            blk = (row0, n, m)
            row0 += n * m
            return blk

        if len(self.o_kf) > 0:
            P, Lm = pose[self.o_kf], lm[self.o_lm]
            E = self.extr_kf[self.o_kf]
            info["reproj"] = add_block(
                lambda d: reproj_residual(pose_plus(P, d[0]), Lm + d[1], E, self.o_intr, self.o_dist, self.o_uv, self.o_sigma,
                                          self.o_cam, self.o_dm, self.o_xi),
                [P, Lm], [self.col_pose[self.o_kf], self.col_lm[self.o_lm]], [6, 3], self.cauchy_reproj)
        if self.imu is not None:
            Pi, Si, Pj, Sj = pose[self.imu_i], sb[self.imu_i], pose[self.imu_j], sb[self.imu_j]
            info["imu"] = add_block(
                lambda d: imu_residual(pose_plus(Pi, d[0]), Si + d[1], pose_plus(Pj, d[2]), Sj + d[3], self.imu),
                [Pi, Si, Pj, Sj], [self.col_pose[self.imu_i], self.col_sb[self.imu_i], self.col_pose[self.imu_j],
                                   self.col_sb[self.imu_j]], [6, 9, 6, 9], None)
        if self.edges is not None:
            e = self.edges
            P1, P2 = pose[e["i"]], pose[e["j"]]
            a2 = torch.where(torch.tensor(e["robust"]), torch.tensor(float(e["a"]) ** 2), torch.tensor(float("inf")))
            info["edges"] = add_block(
                lambda d: between_residual(pose_plus(P1, d[0]), pose_plus(P2, d[1]), e["q"], e["t"], e["S"]),
                [P1, P2], [self.col_pose[e["i"]], self.col_pose[e["j"]]], [6, 6], a2)
        r = torch.cat(res).numpy() if res else np.zeros(0)
        J = None
        if with_jac:
            J = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(row0, self.n)) \
                if vals else sp.csr_matrix((row0, self.n))
        return cost, r, J, info

    def plus(self, pose, sb, lm, delta):
        d = torch.tensor(delta)
        per = self.per
        dc = d[:self.ncam].reshape(self.K, per)
        dp = dc[:, :6].clone()
        dp[torch.tensor(self.const)] = 0
        npose = pose_plus(pose, dp)
        nsb = sb if self.visual_only else sb + dc[:, 6:15]
        nlm = lm.clone()
        ids = np.flatnonzero(self.lm_in)
        if len(ids):
            nlm[ids] = lm[ids] + d[self.ncam:].reshape(-1, 3)
        return npose, nsb, nlm


# ------------------------------------------------------------------------------------------------ linear solve
def solve_normal_equations(J, r, D, ncam):
    """x = argmin |J x - r|^2 + |D x|^2 by eliminating the landmark block (Schur), dense Cholesky of the RCS.
    Zero columns (constant / unused parameters) get x = 0.  Returns None if the factorisation fails."""
    n = J.shape[1]
    H = (J.T @ J).tocsr() + sp.diags(D * D)
    g = J.T @ r
    live = np.asarray(abs(J).sum(0)).reshape(-1) > 0
    x = np.zeros(n)
    cam = np.flatnonzero(live[:ncam]); lmk = ncam + np.flatnonzero(live[ncam:])
    try:
        if len(lmk) == 0:
            Hc = H[cam][:, cam].toarray()
            c = sla.cho_factor(Hc, lower=True)
            x[cam] = sla.cho_solve(c, g[cam])
            return x
        Hcc = H[cam][:, cam]; W = H[cam][:, lmk]; Hll = H[lmk][:, lmk]
        nl = len(lmk) // 3
        B = sp.bsr_matrix(Hll, blocksize=(3, 3))
        B.sort_indices()
        assert B.data.shape[0] == nl and np.array_equal(B.indices, np.arange(nl)), "Hll must be block diagonal"
        Hinv = sp.bsr_matrix((np.linalg.inv(B.data), B.indices, B.indptr), shape=Hll.shape).tocsr()
        S = (Hcc - W @ Hinv @ W.T).toarray()
        gs = g[cam] - W @ (Hinv @ g[lmk])
        c = sla.cho_factor(S, lower=True)
        xc = sla.cho_solve(c, gs)
        x[cam] = xc
        x[lmk] = Hinv @ (g[lmk] - W.T @ xc)
        if not np.all(np.isfinite(x)):
            return None
        return x
    except (np.linalg.LinAlgError, sla.LinAlgError):
        return None


# ------------------------------------------------------------------------------------------------ trust region
def solve(prob: Problem, max_iters: int, log=None):
    """Ceres 1.14 TrustRegionMinimizer + DoglegStrategy(traditional) [A].  Returns dict(pose, sb, lm, cost history,
    iterations).  `iterations` counts successful + unsuccessful steps as Ceres does."""
    pose, sb, lm = prob.pose.clone(), prob.sb.clone(), prob.lm.clone()
    radius, mu = 1e4, 1e-8
    MIN_MU, MAX_MU, MU_INC = 1e-8, 1.0, 10.0
    min_diag, max_diag = 1e-6, 1e32
    cost, r, J, _ = prob.evaluate(pose, sb, lm)
    col_sq = np.asarray(J.multiply(J).sum(0)).reshape(-1)
    scale = 1.0 / (1.0 + np.sqrt(col_sq))          # Jacobi scaling, fixed at x0
    scale[~prob.active] = 0.0
    J = (J @ sp.diags(scale)).tocsr()
    hist = [cost]
    def xnorm(pose, sb, lm):
        v = float((pose[~torch.tensor(prob.const)] ** 2).sum())
        if not prob.visual_only:
            v += float((sb ** 2).sum())
        if prob.L:
            v += float((lm[torch.tensor(prob.lm_in)] ** 2).sum())
        return math.sqrt(v)
    x_norm = xnorm(pose, sb, lm)
    it, reuse, invalid_run = 0, False, 0
    steps = []
    gn = grad = diag = None
    alpha = 0.0
    dogleg_norm = 0.0
    term = "NO_CONVERGENCE"
    g0 = J.T @ r
    if np.max(np.abs(g0 / np.where(scale > 0, scale, 1))) <= 1e-10:
        return dict(pose=pose, sb=sb, lm=lm, cost=hist, iterations=0, termination="CONVERGENCE(gradient)", steps=steps)
    while it < max_iters:
        it += 1
        # ---- DoglegStrategy::ComputeStep ----
        solver_ok = True
        if not reuse:
            reuse = True
            diag = np.sqrt(np.clip(np.asarray(J.multiply(J).sum(0)).reshape(-1), min_diag, max_diag))
            grad = (J.T @ r) / diag
            Jg = J @ (grad / diag)
            alpha = float(grad @ grad) / float(Jg @ Jg)
            gn = None
            while mu < MAX_MU:
                x = solve_normal_equations(J, r, diag * math.sqrt(mu), prob.ncam)
                if x is None:
                    mu *= MU_INC
                    continue
                gn = -x * diag
                break
            solver_ok = gn is not None
        step = None
        if solver_ok:
            gn_norm, g_norm = float(np.linalg.norm(gn)), float(np.linalg.norm(grad))
            if gn_norm <= radius:
                dl, dogleg_norm = gn.copy(), gn_norm
            elif g_norm * alpha >= radius:
                dl, dogleg_norm = -(radius / g_norm) * grad, radius
            else:
                b_dot_a = -alpha * float(grad @ gn)
                a2 = (alpha * g_norm) ** 2
                bma2 = a2 - 2 * b_dot_a + gn_norm ** 2
                c = b_dot_a - a2
                d = math.sqrt(c * c + bma2 * (radius ** 2 - a2))
                beta = (d - c) / bma2 if c <= 0 else (radius ** 2 - a2) / (d + c)
                dl = (-alpha * (1 - beta)) * grad + beta * gn
                dogleg_norm = float(np.linalg.norm(dl))
            step = dl / diag
            step[~prob.active] = 0.0
            mres = J @ step
            model_change = -float(mres @ (r + 0.5 * mres))
        if step is None or model_change <= 0:
            invalid_run += 1
            steps.append(("invalid", None))
            if invalid_run > 5:
                term = "FAILURE(invalid steps)"
                break
            mu *= MU_INC; reuse = False           # StepIsInvalid
            hist.append(cost)
            continue
        invalid_run = 0
        delta = step * scale
        cpose, csb, clm = prob.plus(pose, sb, lm, delta)
        ccost, _, _, _ = prob.evaluate(cpose, csb, clm, with_jac=False)
        step_norm = math.sqrt(float(((cpose - pose) ** 2).sum() + ((csb - sb) ** 2).sum() + ((clm - lm) ** 2).sum()))
        if step_norm <= 1e-8 * (x_norm + 1e-8):
            term = "CONVERGENCE(parameter)"; steps.append(("param_tol", None)); break
        if abs(cost - ccost) <= 1e-6 * cost:
            term = "CONVERGENCE(function)"; steps.append(("func_tol", None)); break
        rho = (cost - ccost) / model_change
        if log:
            log(f"it {it}: cost {cost:.6e} -> {ccost:.6e} rho {rho:.3f} radius {radius:.3e} |step| {step_norm:.3e}")
        if rho > 1e-3:
            pose, sb, lm, cost = cpose, csb, clm, ccost
            x_norm = xnorm(pose, sb, lm)
            _, r, J, _ = prob.evaluate(pose, sb, lm)
            J = (J @ sp.diags(scale)).tocsr()
            if rho < 0.25:
                radius *= 0.5
            if rho > 0.75:
                radius = max(radius, 3.0 * dogleg_norm)
            mu = max(MIN_MU, 2.0 * mu / MU_INC)
            reuse = False
            steps.append(("accepted", rho))
        else:
            radius *= 0.5
            reuse = True
            steps.append(("rejected", rho))
        hist.append(cost)
    return dict(pose=pose, sb=sb, lm=lm, cost=hist, iterations=it, termination=term, steps=steps, radius=radius)


# ------------------------------------------------------------------------------------------------ reference drivers
def corrected_reproj_norms(prob: Problem, pose, sb, lm):
    """problem.Evaluate(residual_ids) of optimization_be.cpp:270-274: loss-corrected residual norms per observation."""
    _, r, _, info = prob.evaluate(pose, sb, lm, with_jac=False)
    r0, n, m = info["reproj"]
    return np.linalg.norm(r[r0:r0 + n * m].reshape(n, m), axis=1)


def global_bundle_adjustment(p, iterations_limit=10, visual_only=False, outlier_removal=True, th_outlier=0.92, log=None):
    """Optimization::GlobalBundleAdjustment (optimization_be.cpp:56-618) on the flat problem.
    Round 1: 5 iterations, loop edges without loss, then erase observations whose corrected residual norm exceeds
    th_gba_outlier_global (0.92, config_backend.yaml:118).  Round 2: `iterations_limit` iterations, loop edges with
    Cauchy(1).  Returns dict(pose, speedbias, lm, obs_removed, lm_included, r1, r2)."""
    p = dict(p)
    n_obs = len(p["obs_kf"])
    removed = np.zeros(n_obs, bool)
    out = {}
    if outlier_removal:
        pr = Problem(p, visual_only=visual_only, loop_loss=None)
        r1 = solve(pr, 5, log)
        norms = corrected_reproj_norms(pr, r1["pose"], r1["sb"], r1["lm"])
        removed[pr.obs_sel[norms > th_outlier]] = True
        out["r1"] = r1
        # round 1 writes nothing back: round 2 restarts from the map state (opt.cpp:325, 454-457), minus the
        # erased observations
    pr2 = Problem(p, visual_only=visual_only, loop_loss=1.0, use_obs=~removed)
    r2 = solve(pr2, iterations_limit, log)
    out.update(pose=r2["pose"].numpy(), speedbias=r2["sb"].numpy(), lm=r2["lm"].numpy(), obs_removed=removed,
               lm_included=pr2.lm_in, r2=r2)
    return out


def pgo_edges(p, vio_pose, wt=(10.0, 1.0, 10.0, 2.0, 3.0), covins_mode=True, use_robust=True, use_nbr=True):
    """Edge list of Optimization::PoseGraphOptimization (optimization_be.cpp:886-1021): loop edges (sqrt_info =
    KF weights in COVINS mode, else chol(cov^-1)^T), successor edges and 5 predecessor edges from the VIO poses
    with weights /1,/2,/2,/3,/3, de-duplicated on the ordered (kf, other) pair."""
    wt_r, wt_t, n1, n23, n45 = wt
    S = np.diag([wt_r] * 3 + [wt_t] * 3) * n1
    Rv = _quat_to_rot_np(vio_pose[:, :4]); tv = vio_pose[:, 4:]
    I, Jj, Q, T, SI, RB = [], [], [], [], [], []
    for l in range(len(p["loop_i"])):
        Sl = S if covins_mode else np.linalg.cholesky(np.linalg.inv(p["loop_cov"][l])).T
        I.append(int(p["loop_i"][l])); Jj.append(int(p["loop_j"][l])); Q.append(p["loop_q"][l]); T.append(p["loop_t"][l])
        SI.append(Sl); RB.append(use_robust)
    seen = set()
    agent, kid = p["agent_of"], p["kf_id"]
    K = int(p["K"])

    def rel(i, j):
        R = Rv[i].T @ Rv[j]
        return _rot_to_quat_np(R), Rv[i].T @ (tv[j] - tv[i])
    for i in range(K):                      # successor edges (:947-972)
        j = i + 1
        if j >= K or agent[j] != agent[i]:
            continue
        if (i, j) in seen:
            continue
        seen.add((i, j))
        q, t = rel(i, j)
        I.append(i); Jj.append(j); Q.append(q); T.append(t); SI.append(S); RB.append(False)
    if use_nbr:
        for i in range(K):                  # 5 predecessors (:976-1021)
            for k in range(1, 6):
                if int(kid[i]) - k > 0:
                    j = i - k
                    Sk = S if k <= 1 else (S / n23 if k <= 3 else S / n45)
                    if (i, j) in seen:
                        continue
                    seen.add((i, j))
                    q, t = rel(i, j)
                    I.append(i); Jj.append(j); Q.append(q); T.append(t); SI.append(Sk); RB.append(False)
    return dict(i=np.array(I, np.int32), j=np.array(Jj, np.int32), q=np.array(Q), t=np.array(T),
                sqrt_info=np.array(SI), robust=np.array(RB))


def _quat_to_rot_np(q):
    return np.stack([_q2R(x) for x in q])


def _rot_to_quat_np(R):
    w = math.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    x = math.copysign(math.sqrt(max(0.0, 1 + R[0, 0] - R[1, 1] - R[2, 2])) / 2, R[2, 1] - R[1, 2])
    y = math.copysign(math.sqrt(max(0.0, 1 - R[0, 0] + R[1, 1] - R[2, 2])) / 2, R[0, 2] - R[2, 0])
    z = math.copysign(math.sqrt(max(0.0, 1 - R[0, 0] - R[1, 1] + R[2, 2])) / 2, R[1, 0] - R[0, 1])
    q = np.array([x, y, z, w])
    return q / np.linalg.norm(q)


def pose_graph_optimization(p, edges, iterations=10, robust_th=0.5, log=None):
    """Optimization::PoseGraphOptimization solve (optimization_be.cpp:1024-1031): poses only, Cauchy(robust_loss_th)
    on the edges flagged robust."""
    pp = dict(K=p["K"], L=0, pose=p["pose"], pose_const=p["pose_const"], extr=p["extr"], cam_of_kf=p.get("cam_of_kf"))
    pr = Problem(pp, visual_only=True, edges=edges, cauchy_edges=robust_th)
    r = solve(pr, iterations, log)
    return dict(pose=r["pose"].numpy(), result=r)
