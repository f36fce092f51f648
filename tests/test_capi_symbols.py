"""CPU: the C-ABI library loads and exports every symbol include/gpb.h declares; the product fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    import gpslam_b200 as gb
    return gb.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "gpb.h")).read()
    names = set(re.findall(r"\b(gpb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 25
    for n in sorted(names):
        assert hasattr(lib, n), "libgpb.so does not export %s" % n


def test_no_cpu_fallback(lib):
    import gpslam_b200 as gb
    if gb.device_count() > 0:
        pytest.skip("GPU present")
    g = gb.Graph(gb.GPB_POSE3, 8, 1)
    g.add_qc_model(np.eye(6))
    g.add_gp_prior(np.arange(7), np.full(7, 0.1))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        g.finalize()
    with pytest.raises(RuntimeError, match="not finalized"):
        g.linearize()


def test_argument_checks(lib):
    import gpslam_b200 as gb
    g = gb.Graph(gb.GPB_POSE3, 8, 1)
    with pytest.raises(RuntimeError):
        g.add_qc_model(-np.eye(6))          # getQc on a non-SPD model is a checked error here (UB in gp/GPutils.cpp:17-19)
    g.add_qc_model(np.eye(6))
    with pytest.raises(RuntimeError):
        g.add_gp_prior([7], [0.1])          # interval out of range
    with pytest.raises(RuntimeError):
        g.add_gp_prior([0], [-1.0])
    with pytest.raises(RuntimeError):
        g.add_interp_range([0], [3], [1.0], [0.1], [0.1], [0.05])
    with pytest.raises(RuntimeError):
        gb.Graph(gb.GPB_LINEAR, 8, 1, dim=5)
