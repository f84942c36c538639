"""COVINS on-disk map format (covins_b200/mapio.py): byte-level checks of the cereal-binary restatement
against hand-assembled expectations (the encodings are pinned by msg_keyframe.hpp:211-285, msg_landmark.hpp:69-73,
map_be.hpp:126-136) and round trips flat problem → directory → flat problem.  PARITY UNPINNED beyond the in-tree
serialisation code: no saved map ships with the reference."""
import os
import struct

import numpy as np
import pytest

from covins_b200 import mapio, synth, synth_map


def test_landmark_bytes_known_answer():
    lm = dict(id=(7, 1), pos_w=[1.0, -2.0, 0.5], observations={(3, 0): 11, (2, 1): 5}, id_reference=(3, 0))
    b = mapio.encode_landmark(lm)
    exp = struct.pack("<QQ", 7, 1)                                   # id: std::pair<size_t,size_t>
    exp += struct.pack("<ii", 3, 1) + struct.pack("<ddd", 1.0, -2.0, 0.5)   # Eigen 3x1: rows, cols, data
    exp += struct.pack("<Q", 2)                                      # std::map size tag
    exp += struct.pack("<QQi", 2, 1, 5) + struct.pack("<QQi", 3, 0, 11)     # items in key order: (2,1) < (3,0)
    exp += struct.pack("<QQ", 3, 0)                                  # id_reference
    assert b == exp
    back = mapio.decode_landmark(b)
    assert back["id"] == (7, 1) and back["observations"] == lm["observations"] and back["id_reference"] == (3, 0)
    assert np.array_equal(back["pos_w"], [1.0, -2.0, 0.5])
    with pytest.raises(ValueError):
        mapio.decode_landmark(b[:-3])
    with pytest.raises(ValueError):
        mapio.decode_landmark(b + b"\0")


def test_eigen_is_column_major_and_cvmat_layout():
    w = mapio._W()
    w.eigen(np.array([[1.0, 2.0], [3.0, 4.0]]))
    assert bytes(w.b) == struct.pack("<ii", 2, 2) + struct.pack("<dddd", 1.0, 3.0, 2.0, 4.0)      # Eigen default storage
    w = mapio._W()
    w.cvmat(np.arange(6, dtype=np.uint8).reshape(2, 3), mapio.CV_8U)
    assert bytes(w.b) == struct.pack("<iiiB", 2, 3, 0, 1) + bytes(range(6))                       # rows, cols, type, continuous, data
    m, t = mapio._R(bytes(w.b)).cvmat()
    assert t == 0 and np.array_equal(m, np.arange(6, dtype=np.uint8).reshape(2, 3))
    w = mapio._W(); w.f64_vec([0.5, 1.5])
    assert bytes(w.b) == struct.pack("<Qdd", 2, 0.5, 1.5)


def test_mapdata_roundtrip():
    T = np.eye(4); T[:3, 3] = [1, 2, 3]
    m = dict(id_map=4, keyframes1=[(10, 0), (11, 1)], keyframes2=[(3, 2), (4, 2)], transforms12=[T, 2 * T], cov=[np.eye(6), 3 * np.eye(6)])
    back = mapio.decode_mapdata(mapio.encode_mapdata(m))
    assert back["id_map"] == 4 and back["keyframes1"] == m["keyframes1"] and back["keyframes2"] == m["keyframes2"]
    assert np.array_equal(back["transforms12"][1], 2 * T) and np.array_equal(back["cov"][1], 3 * np.eye(6))


@pytest.mark.parametrize("cfg", ["tiny", "small"])
def test_flat_problem_roundtrip(tmp_path, cfg):
    p = synth_map.make_config(cfg)
    rng = np.random.default_rng(1)
    desc = rng.integers(0, 256, (len(p["obs_kf"]), 32), dtype=np.uint8)
    d = str(tmp_path / "map")
    mapio.write_map(d, p, descriptors=desc)
    assert len(os.listdir(os.path.join(d, "keyframes"))) == p["K"] and os.path.exists(os.path.join(d, "mapdata.txt"))
    q = mapio.read_map(d)
    n_obs = np.diff(p["lm_obs_ptr"])
    keep = n_obs >= 2                                  # Map::SaveToFile drops landmarks with < 2 observations
    assert q["K"] == p["K"] and q["L"] == int(keep.sum())
    assert np.allclose(q["pose"][:, 4:], p["pose"][:, 4:], rtol=0, atol=1e-15)
    sgn = np.sign(np.sum(q["pose"][:, :4] * p["pose"][:, :4], 1))[:, None]          # q and -q are the same rotation
    assert np.allclose(q["pose"][:, :4] * sgn, p["pose"][:, :4], rtol=0, atol=1e-12)
    for k in ("speedbias", "imu_dt", "imu_acc", "imu_gyr", "imu_acc0", "imu_gyr0", "imu_noise", "intr", "dist", "loop_t"):
        assert np.array_equal(q[k], p[k]), k
    for k in ("imu_i", "imu_j", "imu_ptr", "agent_of", "kf_id", "pose_const", "loop_i", "loop_j", "cam_of_kf"):
        assert np.array_equal(q[k], p[k]), k
    assert np.array_equal(q["lm"], p["lm"][keep])
    obs_keep = np.repeat(keep, n_obs)
    assert np.array_equal(q["obs_kf"], p["obs_kf"][obs_keep]) and np.array_equal(q["obs_uv"], p["obs_uv"][obs_keep])
    assert np.array_equal(q["obs_sigma"], p["obs_sigma"][obs_keep]) and np.array_equal(q["descriptors"], desc[obs_keep])
    assert np.array_equal(q["lm_obs_ptr"], np.concatenate([[0], np.cumsum(n_obs[keep])]))
    assert np.allclose(q["extr"], p["extr"], rtol=0, atol=1e-12)
    # the keyframe descriptor matrices are what the matching calls take: rows = the keyframe's features
    assert sum(len(m) for m in q["kf_descriptors"]) == len(p["obs_kf"])


def test_reread_map_gives_the_same_cost_in_the_oracle(tmp_path):
    """the flat problem read back from disk is the same optimisation problem (oracle cost at the initial state)"""
    from oracle import ba_oracle as bo
    p = synth_map.make_config("tiny")
    d = str(tmp_path / "map")
    mapio.write_map(d, p)
    q = mapio.read_map(d)
    keep = np.diff(p["lm_obs_ptr"]) >= 2
    c0 = bo.Problem(p, visual_only=False, loop_loss=1.0)
    c1 = bo.Problem(q, visual_only=False, loop_loss=1.0)
    f0 = c0.evaluate(c0.pose, c0.sb, c0.lm)[0]; f1 = c1.evaluate(c1.pose, c1.sb, c1.lm)[0]
    assert np.isclose(float(f0), float(f1), rtol=1e-9)


def test_descriptors_add_roundtrip(tmp_path):
    """descriptors_add_ is the feature set the place-recognition k-NN runs on (placerec_gen_be.cpp:82-100): write_map
    stores it per keyframe, read_map returns it as kf_descriptors_add next to kf_descriptors (the DenseMatcher set)."""
    p = synth_map.make_config("tiny")
    rng = np.random.default_rng(2)
    desc = rng.integers(0, 256, (len(p["obs_kf"]), 32), dtype=np.uint8)
    add = [rng.integers(0, 256, (5 + k % 3, 32), dtype=np.uint8) for k in range(p["K"])]
    kp = [rng.random((len(a), 2)).astype(np.float32) for a in add]
    d = str(tmp_path / "map")
    mapio.write_map(d, p, descriptors=desc, descriptors_add=add, keypoints_add=kp)
    q = mapio.read_map(d)
    assert all(np.array_equal(a, b) for a, b in zip(q["kf_descriptors_add"], add))
    assert all(np.array_equal(a, b) for a, b in zip(q["kf_keypoints_add"], kp))
    assert sum(len(m) for m in q["kf_descriptors"]) == len(p["obs_kf"])
