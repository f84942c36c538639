"""oracle/ba_port.cpp (compiled OpenMP CPU port: analytic Jacobians, Schur, tile-sparse BLAS Cholesky, dogleg) against
oracle/ba_oracle.py (independent: torch autograd Jacobians, scipy sparse normal equations, LAPACK).  The port is the
timed CPU baseline of bench.py and the fast oracle of the BASELINE-sized GPU parity tests, so it is pinned here first."""
import numpy as np
import pytest

from covins_b200 import synth_map
from oracle import ba_oracle as bo, ba_port as bp


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-12)


def _same(got, ref, tol=1e-8, lm_tol=1e-6):
    assert got["iterations"] == ref["iterations"]
    assert got["steps"] == [s[0].replace("func_tol", "converged").replace("param_tol", "converged") for s in ref["steps"]]
    assert np.allclose(got["cost"], np.array(ref["cost"]), rtol=1e-7, atol=0)
    assert _rel(got["pose"], ref["pose"].numpy()) < tol and _rel(got["speedbias"], ref["sb"].numpy()) < tol
    assert _rel(got["lm"], ref["lm"].numpy()) < lm_tol


@pytest.mark.parametrize("cfg,visual_only", [("tiny", False), ("tiny", True), ("small", False), ("small", True)])
def test_port_matches_autograd_oracle(cfg, visual_only):
    p = synth_map.make_config(cfg)
    ref = bo.solve(bo.Problem(p, visual_only=visual_only, loop_loss=1.0), 6)
    got = bp.solve(p, 6, visual_only=visual_only)
    _same(got, ref)
    assert bp.lib().blas, "the BLAS-3 tile kernels (scipy's OpenBLAS) must be in use for the timed baseline"


def test_port_constant_poses_skipped_observations_threads():
    p = synth_map.make_config("tiny")
    p["pose_const"] = p["pose_const"].copy(); p["pose_const"][[3, 7, 20]] = 1
    skip = np.zeros(len(p["obs_kf"]), np.uint8); skip[::7] = 1
    ref = bo.solve(bo.Problem(p, visual_only=False, loop_loss=1.0, use_obs=~skip.astype(bool)), 4)
    got = bp.solve(p, 4, obs_skip=skip, threads=1)
    _same(got, ref)
    got3 = bp.solve(p, 4, obs_skip=skip, threads=3)          # reductions differ only by rounding
    assert _rel(got3["pose"], got["pose"]) < 1e-10
    for k in (0, 3, 7, 20):
        assert np.array_equal(got["pose"][k], p["pose"][k])


def test_port_gba_two_rounds_match_oracle():
    p = synth_map.make_config("small")
    ref = bo.global_bundle_adjustment(p, iterations_limit=6)
    got = bp.global_bundle_adjustment(p, iterations_limit=6)
    assert np.array_equal(got["obs_removed"], ref["obs_removed"]) and got["obs_removed"].sum() > 100
    assert got["iterations"] == ref["r2"]["iterations"]
    assert _rel(got["pose"], ref["pose"]) < 1e-8 and _rel(got["speedbias"], ref["speedbias"]) < 1e-8
    well = ref["lm_included"] & (np.abs(ref["lm"]).max(1) < 100.0)
    assert _rel(got["lm"][well], ref["lm"][well]) < 1e-6
    assert np.array_equal(got["lm"][~ref["lm_included"]], p["lm"][~ref["lm_included"]])


def test_port_pgo_matches_oracle():
    p = synth_map.make_map(seed=5, n_agents=3, kf_per_agent=60, n_lm=10, drift_trans=0.01, drift_yaw_deg=0.1)
    edges = bo.pgo_edges(p, p["pose"])
    ref = bo.pose_graph_optimization(p, edges, iterations=10)
    pp = dict(K=p["K"], L=0, pose=p["pose"], pose_const=p["pose_const"], extr=p["extr"], cam_of_kf=p.get("cam_of_kf"))
    got = bp.solve(pp, 10, visual_only=True, cauchy_reproj=0.0, cauchy_edge=0.5, edges=edges)
    assert got["iterations"] == ref["result"]["iterations"]
    assert _rel(got["pose"], ref["pose"]) < 1e-9 and np.allclose(got["cost"], ref["result"]["cost"], rtol=1e-7)
