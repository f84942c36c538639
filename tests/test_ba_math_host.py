"""CPU: the analytic Jacobians of the CUDA cost functions (ba_math.cuh, host+device code) against central finite
differences, compiled for the host with g++ — catches derivation/indexing errors without a GPU."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_device_math_jacobians_vs_finite_differences(tmp_path):
    exe = str(tmp_path / "ba_math_fd")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "ba_math_fd.cpp")])
    out = subprocess.check_output([exe], text=True)
    errs = [float(x) for x in re.findall(r"max jac err ([0-9.eE+-]+)", out)]
    assert len(errs) == 8, out          # 6 camera/distortion instantiations of the reprojection error + between + IMU
    assert out.count("reproj cam") == 6
    assert max(errs) < 1e-6, out
