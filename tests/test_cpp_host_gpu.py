"""The C++ host mirror (include/tbv_b200.hpp: StructuredKStrongest, MapPointNormal, n_scan_normal_reg, OdometryKeyframeFuser over the
C-ABI) exercised from C++ against the oracle's C++ classes: tests/cpp/test_host_mirror.cpp, built by __graft_entry__.build()."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_mirror")


def build_cpp_test() -> str:
    src = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
    deps = [src, os.path.join(ROOT, "include", "tbv_b200.hpp"), os.path.join(ROOT, "include", "tbv_b200.h")]
    if not os.path.exists(BIN) or any(os.path.getmtime(d) > os.path.getmtime(BIN) for d in deps):
        os.makedirs(os.path.dirname(BIN), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                               "-I/usr/local/cuda/include", "-o", BIN, src, "-L" + os.path.join(ROOT, "tbv_slam_public_b200"), "-ltbv_b200",
                               "-L/usr/local/cuda/lib64", "-lcudart",
                               # relative run path: the tree is built in one place and run in another (the GPU box's snapshot)
                               "-Wl,-rpath,$ORIGIN/../../../tbv_slam_public_b200"])
    return BIN


def test_cpp_header_compiles_as_cxx14(tmp_path):
    """The mirror is header-only C++14 (the reference's dialect, cfear_radarodometry/CMakeLists.txt:4): syntax check without linking."""
    tu = tmp_path / "tu.cpp"
    tu.write_text('#include "tbv_b200.hpp"\n'
                  'template class tbv_b200::CeresLeastSquaresT<tbv_b200::DevicePoseGraph>;   // every member of the device-backed optimiser\n'
                  'template class tbv_b200::PointCloudOdometryFuserT<tbv_b200::DeviceOdometryPrimitives>;   // and of the cloud-interface fuser\n'
                  'int main() { return sizeof(tbv_b200::n_scan_normal_reg) > 0 ? 0 : 1; }\n')
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-fsyntax-only", "-I" + os.path.join(ROOT, "include"), str(tu)])


@pytest.mark.gpu
def test_cpp_host_mirror_against_oracle(tmp_path):
    from tbv_slam_public_b200 import synth
    st = synth.make_stream(4)
    p = tmp_path / "scans.bin"
    np.ascontiguousarray(st.scans).tofile(p)
    exe = build_cpp_test()
    r = subprocess.run([exe, str(p), "4", "400", "3768"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
