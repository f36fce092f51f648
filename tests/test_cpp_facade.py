"""C++ host facade (include/gpslam_b200/gpslam.h): compiles against the C ABI with plain g++; on a GPU box the ported reference
tests (tests/cpp/test_facade.cpp) must pass, on a CPU box the binary must fail loudly (no CPU fallback)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "test_facade")


def _build():
    import __graft_entry__ as ge
    ge.build()
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_facade.cpp"),
                    "-L" + os.path.join(ROOT, "gpslam_b200"), "-lgpb", "-Wl,-rpath," + os.path.join(ROOT, "gpslam_b200"), "-o", BIN], check=True)


def test_facade_compiles_and_fails_loudly_without_gpu():
    import gpslam_b200 as gb
    _build()
    if gb.device_count() > 0:
        pytest.skip("GPU present")
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 2 and "no CUDA device" in r.stdout


@pytest.mark.gpu
def test_facade_reference_style_tests():
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
