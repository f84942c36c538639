"""f1 parity pinned by REFERENCE-WRITTEN bytes: tests/golden/covins_map_ref/ was written by oracle/_ref/cereal_fixture_gen
— the vendored cereal::BinaryOutputArchive + the reference's own message types and `save` templates
(msg_keyframe.hpp:129-143,211-285, msg_landmark.hpp:69-73, typedefs_base.hpp:376-380; generator
oracle/ref/cereal_fixture_gen.cpp, recipe oracle/ref/Makefile).  covins_b200.mapio must (a) decode every file to the
values of manifest.json and (b) re-encode the decoded message to the identical bytes."""
import json
import os
import subprocess

import numpy as np
import pytest

from covins_b200 import mapio

D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "covins_map_ref")
MAN = json.load(open(os.path.join(D, "manifest.json")))


def _rd(*parts):
    with open(os.path.join(D, *parts), "rb") as f:
        return f.read()


def _col(v):
    return np.asarray(v, np.float64).reshape(-1)


@pytest.mark.parametrize("k", [0, 1, 2])
def test_keyframe_bytes_written_by_cereal(k):
    b = _rd("keyframes", f"keyframes{k}.txt")
    kf = mapio.decode_keyframe(b)
    m = MAN["keyframes"][k]
    assert kf["timestamp"] == m["timestamp"] and kf["id"] == tuple(m["id"])
    c = kf["calibration"]
    assert np.array_equal(c["T_SC"], np.array(m["T_SC"])) and c["cam_model"] == m["cam_model"] and c["dist_model"] == m["dist_model"]
    for key in ("img_dims", "dist_coeffs", "intrinsics", "a0"):
        assert np.array_equal(c[key], _col(m[key])), key
    assert np.array_equal(c["K"], np.array(m["K"]))
    for key in ("a_max", "g_max", "sigma_a_c", "sigma_g_c", "sigma_ba", "sigma_bg", "sigma_aw_c", "sigma_gw_c", "tau", "g", "rate",
                "delay_cam0_to_imu", "delay_cam1_to_imu"):
        assert c[key] == m[key], key
    for key in ("img_dim_x_min", "img_dim_y_min", "img_dim_x_max", "img_dim_y_max"):
        assert kf[key] == m[key]
    for sfx in ("", "_add"):
        for key, n in (("keypoints_distorted", 2), ("keypoints_undistorted", 2), ("keypoints_aors", 4)):
            assert np.array_equal(kf[key + sfx], np.array(m[key + sfx], np.float32).reshape(-1, n)), key + sfx
        dm = m["descriptors" + sfx]
        dt = np.float32 if (dm["type"] & 7) == 5 else np.uint8
        assert kf["descriptors" + sfx].dtype == dt
        assert np.array_equal(kf["descriptors" + sfx], np.array(dm["data"], dt).reshape(dm["rows"], dm["cols"]))
    for key in ("T_s_c", "T_w_s", "T_w_s_vio"):
        assert np.array_equal(kf[key], np.array(m[key])), key
    for key in ("velocity", "bias_gyro", "bias_accel", "lin_acc", "ang_vel", "lin_acc_init", "ang_vel_init"):
        assert np.array_equal(kf[key], _col(m[key])), key
    pre = kf["preintegration"]
    for key in ("acc", "gyr", "lin_bias_accel", "lin_bias_gyro"):
        assert np.array_equal(pre[key], _col(m["pre_" + key]))
    for key in ("dt", "lin_acc_x", "lin_acc_y", "lin_acc_z", "ang_vel_x", "ang_vel_y", "ang_vel_z"):
        assert np.array_equal(pre[key], np.array(m["pre_" + key], np.float64))
    assert kf["landmarks"] == {fi: (a, b2) for fi, a, b2 in m["landmarks"]}
    assert kf["id_predecessor"] == tuple(m["id_predecessor"]) and kf["id_successor"] == tuple(m["id_successor"])
    if k != 1:
        assert mapio.DEFPAIR in (kf["id_predecessor"], kf["id_successor"])     # defpair = (KFRANGE, MAPRANGE), typedefs_base.hpp:56
    assert kf["img"].size == 0
    assert mapio.encode_keyframe(kf) == b                                       # byte-for-byte


@pytest.mark.parametrize("l", [0, 1, 2, 3])
def test_landmark_bytes_written_by_cereal(l):
    b = _rd("mappoints", f"mappoints{l}.txt")
    lm = mapio.decode_landmark(b)
    m = MAN["landmarks"][l]
    assert lm["id"] == tuple(m["id"]) and np.array_equal(lm["pos_w"], _col(m["pos_w"]))
    assert lm["observations"] == {(a, c): f for a, c, f in m["observations"]} and lm["id_reference"] == tuple(m["id_reference"])
    assert mapio.encode_landmark(lm) == b


def test_mapdata_bytes_written_by_cereal():
    b = _rd("mapdata.txt")
    md = mapio.decode_mapdata(b)
    m = MAN["mapdata"]
    assert md["id_map"] == m["id_map"]
    assert md["keyframes1"] == [tuple(x) for x in m["keyframes1"]] and md["keyframes2"] == [tuple(x) for x in m["keyframes2"]]
    for e in range(2):
        assert np.array_equal(md["transforms12"][e], np.array(m["transforms12"][e]).reshape(4, 4))
        assert np.array_equal(md["cov"][e], np.array(m["cov"][e]).reshape(6, 6))
    assert mapio.encode_mapdata(md) == b


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(D), "..", "..", "oracle", "_ref", "cereal_fixture_gen")),
                    reason="oracle/_ref not built (needs /root/reference)")
def test_fixture_is_what_the_reference_stack_writes_today(tmp_path):
    gen = os.path.join(os.path.dirname(D), "..", "..", "oracle", "_ref", "cereal_fixture_gen")
    subprocess.check_call([gen, str(tmp_path / "m")])
    for sub, names in (("keyframes", [f"keyframes{k}.txt" for k in range(3)]), ("mappoints", [f"mappoints{l}.txt" for l in range(4)]), ("", ["mapdata.txt"])):
        for n in names:
            assert open(tmp_path / "m" / sub / n, "rb").read() == _rd(sub, n) if sub else open(tmp_path / "m" / n, "rb").read() == _rd(n)
