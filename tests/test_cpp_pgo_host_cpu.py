"""The C++ mirror's pose-graph optimiser (tbv_b200::CeresLeastSquaresT, include/tbv_b200.hpp) on CPU: its trust-region loop with the
oracle's assembly and a dense Cholesky standing in for the two device calls (tests/cpp/test_pgo_host.cpp)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "_build", "test_pgo_host")


def test_cpp_pose_graph_trust_region_loop():
    src = os.path.join(ROOT, "tests", "cpp", "test_pgo_host.cpp")
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    # no -ltbv_b200: the loop is a template over its backend, and this test never instantiates the device backend
    subprocess.check_call(["g++", "-std=c++14", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "oracle"),
                           "-o", BIN, src])
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("PASS"), r.stdout[-2000:] + r.stderr[-2000:]
