"""oracle/relpose_oracle.py — CPU restatement of Optimization::OptimizeRelativePose (optimization_be.cpp:620-831).
TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (robopt_open / Ceres absent, see ba_oracle.py): assumptions [A]

  parameter block     ceresAB = [qx,qy,qz,qw, x,y,z] of T12 (= T_AB, :629-638) with PoseQuaternionLocalParameterization
  residual A [A]      robopt::reprojection::RelativeEuclideanReprError<Cam,Dist>(kpA, sigmaA, camA, P3DBc, kNormal):
                      r = (project_A(R_AB p_B + t_AB) - kpA) / sigmaA                                           (:674-713)
  residual B [A]      ... (kpB, sigmaB, camB, P3DAc, kInverse): r = (project_B(R_AB^T (p_A - t_AB)) - kpB) / sigmaB    (:716-760)
  loss                one shared ceres::CauchyLoss(1.0) on every block (:625-626, :763-770)
  solve               5 iterations (dogleg, Ceres defaults), outlier purge on the loss-corrected residual norms
                      (problem.Evaluate applies the loss) > th_outlier_align of EITHER block (:798-818), < 12 survivors →
                      return 0 with T12 untouched (:821-823), 5 more iterations, T12 = Ceres2Transform (:829)
Jacobians come from torch autograd; the trust-region loop is ba_oracle.solve (the same restated Ceres minimiser)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import torch

from . import ba_oracle as bo


class RelPoseProblem:
    """duck-types what ba_oracle.solve needs: one free pose block (6 local dofs), no speed-bias, no landmarks"""

    def __init__(self, T12, pA_c, pB_c, kpA, kpB, sigmaA, sigmaB, camA, camB, active=None):
        self.K, self.L, self.per, self.ncam, self.n = 1, 0, 6, 6, 6
        self.visual_only = True
        self.pose = torch.tensor(np.asarray(T12, float)).reshape(1, 7)
        self.sb = torch.zeros(1, 9); self.lm = torch.zeros(0, 3)
        self.const = np.array([False]); self.lm_in = np.zeros(0, bool); self.active = np.ones(6, bool)
        self.pA, self.pB = torch.tensor(np.asarray(pA_c, float)), torch.tensor(np.asarray(pB_c, float))
        self.kpA, self.kpB = torch.tensor(np.asarray(kpA, float)), torch.tensor(np.asarray(kpB, float))
        self.sA, self.sB = torch.tensor(np.asarray(sigmaA, float)), torch.tensor(np.asarray(sigmaB, float))
        self.camA, self.camB = camA, camB
        n = len(self.pA)
        self.keep = np.ones(n, bool) if active is None else np.asarray(active, bool)

    @staticmethod
    def _project(cam, pc, kp, sigma):
        n = pc.shape[0]
        ident = torch.zeros(n, 7); ident[:, 3] = 1.0                      # identity pose / extrinsics: pc is already in the camera
        intr = torch.tensor(np.asarray(cam["intr"], float)).expand(n, 4); dist = torch.tensor(np.asarray(cam["dist"], float)).expand(n, 4)
        cm = torch.full((n,), int(cam.get("cam_model", 0)), dtype=torch.long); dm = torch.full((n,), int(cam.get("dist_model", 0)), dtype=torch.long)
        xi = torch.full((n,), float(cam.get("xi", 0.0)))
        return bo.reproj_residual(ident, pc, ident, intr, dist, kp, sigma, cm, dm, xi)

    def residuals(self, pose):
        q, t = pose[:, :4], pose[:, 4:]
        sel = torch.tensor(self.keep)
        pa = bo.qrot(q, self.pB[sel]) + t                                  # kNormal: B's point into camera A
        pb = bo.qrot(bo.qconj(q), self.pA[sel] - t)                        # kInverse: A's point into camera B
        return self._project(self.camA, pa, self.kpA[sel], self.sA[sel]), self._project(self.camB, pb, self.kpB[sel], self.sB[sel])

    def evaluate(self, pose, sb, lm, with_jac=True):
        d = torch.zeros(1, 6, requires_grad=with_jac)
        rA, rB = self.residuals(bo.pose_plus(pose, d))
        r = torch.cat([rA, rB], 0)                                         # block order A_0..A_n, B_0..B_n (irrelevant for the sums)
        s = (r.detach() ** 2).sum(1)
        scale = torch.sqrt(1.0 / (1.0 + s))                                # Cauchy(1) corrector
        cost = float((0.5 * torch.log1p(s)).sum())
        res = (r.detach() * scale[:, None]).reshape(-1).numpy()
        J = None
        if with_jac:
            rows = []
            for k in range(r.shape[0]):
                for c in range(2):
                    g, = torch.autograd.grad(r[k, c], d, retain_graph=True)
                    rows.append((g[0] * scale[k]).numpy())
            J = sp.csr_matrix(np.array(rows).reshape(-1, 6))
        return cost, res, J, {"n": int(rA.shape[0])}

    def plus(self, pose, sb, lm, delta):
        return bo.pose_plus(pose, torch.tensor(delta).reshape(1, 6)), sb, lm

    def corrected_norms(self, pose):
        rA, rB = self.residuals(pose)
        f = lambda r: (r.norm(dim=1) * torch.sqrt(1.0 / (1.0 + (r ** 2).sum(1)))).numpy()
        return f(rA), f(rB)


def optimize_relative_pose(T12, pA_c, pB_c, kpA, kpB, sigmaA, sigmaB, camA, camB, th_outlier_align=1.3):
    """→ dict(T12 [7] (input when the return value is 0), removed [n] bool (by residual index), n_inliers, r1, r2)"""
    n = len(pA_c)
    pr = RelPoseProblem(T12, pA_c, pB_c, kpA, kpB, sigmaA, sigmaB, camA, camB)
    r1 = bo.solve(pr, 5)
    nA, nB = pr.corrected_norms(r1["pose"])
    removed = (nA > th_outlier_align) | (nB > th_outlier_align)
    out = dict(removed=removed, r1=r1)
    if n - int(removed.sum()) < 12:
        out.update(T12=np.asarray(T12, float).copy(), n_inliers=0, r2=None)
        return out
    pr2 = RelPoseProblem(r1["pose"].numpy().reshape(7), pA_c, pB_c, kpA, kpB, sigmaA, sigmaB, camA, camB, active=~removed)
    r2 = bo.solve(pr2, 5)
    q = r2["pose"].numpy().reshape(7).copy()
    q[:4] /= np.linalg.norm(q[:4])                                          # Utils::Ceres2Transform normalises (utils_base.cpp:38-40)
    out.update(T12=q, n_inliers=n - int(removed.sum()), r2=r2)
    return out
