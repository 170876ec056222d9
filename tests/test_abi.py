"""CPU tier: libb2sense.so loads, exports every symbol include/b2sense.h declares, validates its
arguments, and FAILS LOUDLY without a GPU (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mujoco_ros2_simulation_b200 import _capi, engine, scenes

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, "include", "b2sense.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = C.CDLL(_capi.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in b2sense.h but not exported"
    assert declared == set(_capi.SYMBOLS), declared ^ set(_capi.SYMBOLS)


def test_version_and_device_count():
    lib = _capi.load()
    assert lib.b2_version() == 100
    assert lib.b2_device_count() >= 0


def test_library_is_sm100a_only():
    """the shared object carries sm_100a SASS (cuobjdump) and nothing else"""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_argument_validation():
    lib = _capi.load()
    h = C.c_void_p()
    assert lib.b2_scene_create(None, 1, 0, None, C.byref(h)) == _capi.B2_ERR_INVALID
    assert b"null" in lib.b2_last_error()
    rs = scenes.robot_scene()
    desc = engine._model_desc(rs.model)
    assert lib.b2_scene_create(C.byref(desc), 0, 0, None, C.byref(h)) == _capi.B2_ERR_INVALID
    bad = engine._model_desc(rs.model)
    t = rs.model.geom_type.copy()
    t[0] = 1  # hfield
    bad.geom_type = t.ctypes.data_as(_capi.i32p)
    assert lib.b2_scene_create(C.byref(bad), 1, 0, None, C.byref(h)) == _capi.B2_ERR_UNSUPPORTED
    assert lib.b2_scene_update_poses(None, 0, 1, None, None) == _capi.B2_ERR_INVALID
    assert lib.b2_lidar_cast(None, 1, None, None, None, None, None, None, None) == _capi.B2_ERR_INVALID
    assert lib.b2_multiray(None, 0, None, None, None, 1, -1, None, None, 1, 1.0) == _capi.B2_ERR_INVALID
    cd = _capi.CameraDesc(640, 480, 58.0, 0.01, 50.0, 1.0, 13)
    assert lib.b2_camera_create(None, 0, C.byref(cd), C.byref(h)) == _capi.B2_ERR_INVALID
    assert lib.b2_scene_destroy(None) == 0 and lib.b2_lidar_destroy(None) == 0


def test_no_cpu_fallback(has_gpu):
    """without a CUDA device the product refuses to work instead of computing on the CPU"""
    if has_gpu:
        pytest.skip("a GPU is present")
    rs = scenes.robot_scene()
    with pytest.raises(_capi.B2Error) as e:
        engine.Scene(rs.model)
    assert e.value.code == _capi.B2_ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(_capi.B2Error):
        engine.DepthCamera(scenes.depth_camera())


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under the package may reference it"""
    pkg = os.path.join(ROOT, "mujoco_ros2_simulation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f)).read()
                assert "b2_oracle" not in src and "oracle.binding" not in src and "from oracle" not in src, f
                assert "emu_binding" not in src, f
