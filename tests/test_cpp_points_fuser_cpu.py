"""The C++ mirror's cloud-interface fuser (tbv_b200::PointCloudOdometryFuserT, include/tbv_b200.hpp) on CPU: processFrame's bookkeeping with
the oracle's primitives as backend, against the oracle's fused frame on the same scans (tests/cpp/test_points_fuser.cpp)."""
import os
import subprocess

import numpy as np

from tbv_slam_public_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_points_fuser")


def test_cpp_cloud_interface_fuser_equals_the_fused_frame(tmp_path):
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-pthread", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                           "-o", BIN, os.path.join(ROOT, "tests", "cpp", "test_points_fuser.cpp")])
    st = synth.make_stream(14, speed=4.0)
    p = tmp_path / "scans.bin"
    np.ascontiguousarray(st.scans).tofile(p)
    r = subprocess.run([BIN, str(p), "14", "400", "3768"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("PASS"), r.stdout[-2000:] + r.stderr[-2000:]
