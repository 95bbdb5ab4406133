"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/rspcl.h declares, the
struct layouts the bindings assume match the header, and the product fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import rspcl_b200 as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "rspcl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rspcl_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = R.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(L, s), "librspcl_b200.so does not export %s" % s
    assert sorted(R.EXPORTS) == syms


def test_struct_layouts_match_header():
    assert C.sizeof(R.IcpParams) == 40 and C.sizeof(R.IcpResult) == 96
    assert C.sizeof(R.NdtParams) == 48 and C.sizeof(R.NdtResult) == 144
    assert C.sizeof(R.PairResult) == 160
    assert R.POINT.itemsize == 16 and R.PCL32.itemsize == 32 and R.NDT_VOXEL.itemsize == 224


def test_reference_parameter_defaults():
    p = R.icp_params()   # icp_edge_based_registration.hpp:42-45
    assert (p.max_iterations, p.max_corr_dist, p.transformation_epsilon, p.euclidean_fitness_epsilon) == (100, 0.01, 1.0, 1000.0)
    assert p.min_correspondences == 3 and p.mse_threshold_absolute == 1e-12
    q = R.ndt_params()   # ndt_edge_based_registration.hpp:39-43
    assert (q.max_iterations, q.transformation_epsilon, q.step_size, q.resolution) == (50, 0.01, 0.1, 1.0)
    assert (q.min_points_per_voxel, q.outlier_ratio, q.min_covar_eigvalue_mult) == (6, 0.55, 0.01)


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(R.RspclError):
        R.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "realsense-pointcloud_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "liborc" not in txt and "import orc" not in txt and "oracle/" not in txt.replace("oracle/orc_edge.cpp", ""), \
                    "%s references the oracle" % os.path.join(dp, f)


def test_pcl32_layout_helpers_round_trip():
    rng = np.random.default_rng(0)
    p = np.zeros(100, R.POINT)
    p["x"], p["y"], p["z"] = rng.normal(size=(3, 100)).astype(np.float32)
    p["rgba"] = rng.integers(0, 2**32, 100, dtype=np.uint64).astype(np.uint32)
    q = R.to_pcl32(p)
    assert q.dtype.itemsize == 32 and (q["w"] == 1.0).all()
    assert np.array_equal(R.from_pcl32(q), p)
