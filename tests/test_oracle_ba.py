"""CPU: invariants that pin the BA/PGO oracle itself (it is "parity unpinned" w.r.t. Ceres/robopt — see its header):
finite differences of the residual functions vs the autograd Jacobians in the LOCAL parametrisation, zero residual on
noise-free data, monotone cost over accepted steps, gauge, IMU preintegration consistency."""
import numpy as np
import pytest
import torch

from covins_b200 import synth_map
from oracle import ba_oracle as bo


@pytest.fixture(scope="module")
def tiny():
    return synth_map.make_config("tiny")


def _fd_jac(prob, pose, sb, lm, eps=1e-6):
    _, r0, _, _ = prob.evaluate(pose, sb, lm, with_jac=False)
    cols = []
    idx = np.flatnonzero(prob.active)
    rng = np.random.default_rng(0)
    pick = rng.choice(idx, size=min(40, len(idx)), replace=False)
    J = np.zeros((len(r0), len(pick)))
    for n, c in enumerate(pick):
        d = np.zeros(prob.n); d[c] = eps
        rp = prob.evaluate(*prob.plus(pose, sb, lm, d), with_jac=False)[1]
        d[c] = -eps
        rm = prob.evaluate(*prob.plus(pose, sb, lm, d), with_jac=False)[1]
        J[:, n] = (rp - rm) / (2 * eps)
    return pick, J


@pytest.mark.parametrize("visual_only", [True, False])
def test_autograd_jacobian_matches_finite_differences(tiny, visual_only):
    # no robust loss here: the corrector's Jacobian is by design NOT the derivative of the corrected residual
    prob = bo.Problem(tiny, visual_only=visual_only, cauchy_reproj=None, loop_loss=None)
    _, r, J, _ = prob.evaluate(prob.pose, prob.sb, prob.lm)
    pick, Jfd = _fd_jac(prob, prob.pose, prob.sb, prob.lm)
    Ja = J[:, pick].toarray()
    scale = np.maximum(np.abs(Ja).max(0), 1.0)
    assert np.abs(Ja - Jfd).max(0).max() / scale.max() < 1e-5
    assert (np.abs(Ja - Jfd) / scale[None, :]).max() < 1e-5


def test_zero_residual_at_ground_truth_without_noise():
    p = synth_map.make_map(seed=11, n_agents=1, kf_per_agent=10, n_lm=150, pix_noise=0.0, outlier_frac=0.0, with_imu=False)
    p["pose"] = p["gt_pose"].copy(); p["lm"] = p["gt_lm"].copy()
    prob = bo.Problem(p, visual_only=True, loop_loss=None)
    cost, r, _, info = prob.evaluate(prob.pose, prob.sb, prob.lm, with_jac=False)
    r0, n, m = info["reproj"]
    assert np.abs(r[r0:r0 + n * m]).max() < 1e-4   # obs_uv is float32
    res = bo.solve(prob, 5)
    assert res["cost"][-1] <= res["cost"][0] + 1e-12


def test_imu_preintegration_matches_ground_truth_motion(tiny):
    """VINS-style residual at ground-truth states (true biases) must be a few sigma, not thousands."""
    p = dict(tiny)
    p["pose"] = p["gt_pose"]; p["speedbias"] = p["gt_speedbias"]
    prob = bo.Problem(p, visual_only=False, loop_loss=None)
    _, r, _, info = prob.evaluate(prob.pose, prob.sb, prob.lm, with_jac=False)
    r0, n, m = info["imu"]
    chi = (r[r0:r0 + n * m].reshape(n, m) ** 2).sum(1)
    assert np.median(chi) < 60.0, np.median(chi)    # 15 dof, midpoint discretisation adds a little


@pytest.mark.parametrize("visual_only", [True, False])
def test_cost_monotone_and_gauge_fixed(tiny, visual_only):
    prob = bo.Problem(tiny, visual_only=visual_only, loop_loss=1.0)
    res = bo.solve(prob, 6)
    c = res["cost"]
    assert all(c[i + 1] <= c[i] * (1 + 1e-12) for i in range(len(c) - 1))
    assert c[-1] < c[0]
    assert torch.equal(res["pose"][0], prob.pose[0])          # KF (0, map id) is constant (opt.cpp:88-89)
    assert res["iterations"] <= 6


def test_gba_two_rounds_remove_gross_outliers():
    p = synth_map.make_config("small")
    out = bo.global_bundle_adjustment(p, iterations_limit=4, visual_only=True)
    rem, truth = out["obs_removed"], p["obs_is_outlier"]
    assert rem.sum() > 0.5 * truth.sum()
    assert (rem & truth).sum() / max(rem.sum(), 1) > 0.8        # what round 1 erases is mostly the injected 5 %


def test_pgo_pulls_drifted_trajectory_towards_loop_constraints():
    p = synth_map.make_map(seed=5, n_agents=2, kf_per_agent=30, n_lm=10, drift_trans=0.01, drift_yaw_deg=0.1)
    edges = bo.pgo_edges(p, p["pose"])
    assert len(edges["i"]) > 5 * p["K"]
    res = bo.pose_graph_optimization(p, edges, iterations=10)
    r = res["result"]
    assert r["cost"][-1] < 0.8 * r["cost"][0]


def test_product_pgo_edge_builder_equals_oracle():
    """host logic of PoseGraphOptimization (edge construction, optimization_be.cpp:886-1021): the vectorised builder in
    covins_b200.optimization must produce the oracle's edge list (same order, weights, robust flags)."""
    from covins_b200 import optimization as O
    from covins_b200 import synth_map
    for covins_mode in (True, False):
        p = synth_map.make_map(seed=6, n_agents=3, kf_per_agent=40, n_lm=10, drift_trans=0.01, drift_yaw_deg=0.1)
        ref = bo.pgo_edges(p, p["pose"], covins_mode=covins_mode)
        got = O.pgo_edges(p, p["pose"], covins_mode=covins_mode)
        assert np.array_equal(got["i"], ref["i"]) and np.array_equal(got["j"], ref["j"])
        assert np.array_equal(got["robust"], ref["robust"])
        assert np.allclose(got["sqrt_info"], ref["sqrt_info"], rtol=1e-14, atol=0)
        # small quaternion components come out of sqrt(1 + R00 - R11 - R22): rounding of R is amplified to ~1e-12 there
        assert np.allclose(got["q"], ref["q"], rtol=0, atol=5e-11) and np.allclose(got["t"], ref["t"], rtol=0, atol=1e-13)


@pytest.mark.parametrize("cam_model,dist_model", [(0, 1), (0, 2), (1, 0), (1, 1), (1, 2)])
def test_oracle_camera_models_reduce_cost_to_noise_level(cam_model, dist_model):
    """the other five GlobalEuclideanReprError instantiations (optimization_be.cpp:186-231) in the autograd oracle: a map
    observed through the model is solved back to the noise level (inlier residuals ~ 1 px / sigma)"""
    from covins_b200 import synth_map
    p = synth_map.with_camera_model(synth_map.make_config("tiny"), cam_model, dist_model, xi=0.9 if cam_model else 0.0, seed=7)
    pr = bo.Problem(p, visual_only=True, loop_loss=1.0)
    r = bo.solve(pr, 6)
    assert r["cost"][-1] < r["cost"][0]
    norms = bo.corrected_reproj_norms(pr, r["pose"], r["sb"], r["lm"])
    inl = ~p["obs_is_outlier"][pr.obs_sel]
    assert np.median(norms[inl]) < 0.5
