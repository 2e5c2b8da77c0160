"""The C-ABI library loads, exports every symbol include/b200slam.h declares, validates arguments
without a device and fails loudly (no CPU fallback) when there is none."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import helpers as H
from slam_toolbox_b200 import api, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b200slam.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b200[a-z_0-9]*)\s*\(", txt)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    L = C.CDLL(path)
    syms = declared_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_kernels_are_compiled_for_sm_100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_argument_validation_needs_no_device():
    L = api.lib()
    h = C.c_void_p()
    p = api.SmParams(1.0, 0.0, 0.03, 12.0, 0.3, 0.03, 0.003, 0.25, 1.0, 0.5, 0.9, 0)   # resolution 0
    assert L.b200sm_create(C.byref(p), C.byref(h)) == api.ERR_INVALID_ARG
    assert b"Mapper.cpp:481" in L.b200_last_error()
    p = api.SmParams(1.0, 0.05, 1.0, 12.0, 0.3, 0.03, 0.003, 0.25, 1.0, 0.5, 0.9, 0)   # smear too large
    assert L.b200sm_create(C.byref(p), C.byref(h)) == api.ERR_INVALID_ARG
    assert L.b200sm_create(None, C.byref(h)) == api.ERR_INVALID_ARG
    assert L.b200sm_match(None, None, None, 0, 0, 0, None, None, None) == api.ERR_INVALID_ARG
    assert L.b200pg_add_node(None, 0, None) == api.ERR_INVALID_ARG
    og = api.OgParams()
    L.b200og_default_params(C.byref(og))
    assert (og.resolution, og.min_pass_through, og.occupancy_threshold) == (0.05, 2, 0.1)
    og.resolution = 0.0                                                              # "Resolution cannot be 0", Karto.h:5916-5918
    assert L.b200og_create(C.byref(og), C.byref(h)) == api.ERR_INVALID_ARG and b"Karto.h:5916" in L.b200_last_error()
    assert L.b200og_create(None, C.byref(h)) == api.ERR_INVALID_ARG
    assert L.b200og_add_scans(None, None, 0) == api.ERR_INVALID_ARG and L.b200og_build(None, None) == api.ERR_INVALID_ARG
    assert L.b200og_num_scans(None) == 0
    o = api.PgOpts()
    L.b200pg_default_opts(C.byref(o))
    assert (o.max_num_iterations, o.function_tolerance, o.initial_trust_region_radius, o.jacobi_scaling) == (50, 1e-3, 1e4, 1)


def test_point_readings_match_the_oracle():
    from oracle import karto_port as P
    from slam_toolbox_b200 import synth
    rng = np.random.default_rng(0)
    r = rng.uniform(0.1, 30.0, 1081)
    r[::50] = np.inf
    pose = np.array([3.25, -7.5, 0.7])
    a = api.point_readings(r, pose, api.LaserRangeFinder())[0]
    b = P.point_readings(r, pose, synth.ANGLE_MIN, synth.ANGLE_INC)
    assert np.array_equal(a, b, equal_nan=True)


@pytest.mark.skipif(H is None, reason="")
def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(api.B200Error) as e:
        H.gpu_matcher(H.MAPPER_LOOP, H.GRID_SMALL)
    assert e.value.code == api.ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(api.B200Error) as e:
        api.ScanSolver()
    assert e.value.code == api.ERR_CUDA
    with pytest.raises(api.B200Error) as e:
        api.OccupancyGrid(0.05)
    assert e.value.code == api.ERR_CUDA and "no CPU fallback" in str(e.value)
