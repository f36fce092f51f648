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


def test_facade_archives_round_trip():
    """include/gpslam_b200/archive.h + the serialize() members of every facade class (the reference's classes are
    boost-serializable): object -> text -> object for every factor class, interpolator, value type, a graph and Values; type
    restoration behind the base pointer, shared noise models, files, and loud failure on damaged input.  No device call: runs here."""
    import __graft_entry__ as ge
    ge.build()
    exe = os.path.join(ROOT, "tests", "cpp", "test_archive")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_archive.cpp"),
                        "-L" + os.path.join(ROOT, "gpslam_b200"), "-lgpb", "-Wl,-rpath," + os.path.join(ROOT, "gpslam_b200"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "archive tests passed" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_facade_reference_style_tests():
    _build()
    r = subprocess.run([BIN], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_gtsam_adapter_lowering_optimises_on_gpu():
    """the same adapter run, in the driver's GPU tier (the unmarked twin below is deselected by -m gpu)"""
    test_gtsam_adapter_compiles_against_api_stubs()


def test_gtsam_adapter_compiles_against_api_stubs():
    """include/gpslam_b200/gtsam_adapter.h cannot meet the real GTSAM here (absent: SURVEY.md §8c); it is
    type-checked, warning-free, against tests/cpp/gtsam_stub (declarations of the GTSAM / gpslam entry points it calls) and its
    lowering is run on a small Plaza-shaped graph: without a GPU the run ends at gpb_graph_finalize (loudly), with one it optimises"""
    import __graft_entry__ as ge
    import gpslam_b200 as gb
    ge.build()
    exe = os.path.join(ROOT, "tests", "cpp", "test_gtsam_adapter")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "tests", "cpp", "gtsam_stub"), "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "cpp", "test_gtsam_adapter.cpp"), "-L" + os.path.join(ROOT, "gpslam_b200"), "-lgpb",
                        "-Wl,-rpath," + os.path.join(ROOT, "gpslam_b200"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    if gb.device_count() > 0:
        assert r.returncode == 0 and "optimised" in r.stdout, r.stdout + r.stderr
    else:
        assert r.returncode == 2 and "no CUDA device" in r.stdout, r.stdout + r.stderr
