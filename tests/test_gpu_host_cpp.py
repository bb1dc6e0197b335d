"""Builds and runs the C++ host-mirror parity test (tests/cpp/test_host_mirror.cpp) on the GPU."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_host_mirror_against_oracle(tmp_path, orc):
    exe = str(tmp_path / "test_host_mirror")
    pkg = os.path.join(ROOT, "interactive-rate-tendons_b200")
    cmd = ["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp"), "-o", exe,
           "-L" + pkg, "-lirt_b200", "-L" + os.path.join(ROOT, "oracle"), "-loracle",
           "-Wl,-rpath," + pkg, "-Wl,-rpath," + os.path.join(ROOT, "oracle"), "-fopenmp"]
    subprocess.check_call(cmd, env=dict(os.environ, CXX="g++"))
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(out.stdout[-3000:], out.stderr[-2000:])
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "host mirror ok" in out.stdout
