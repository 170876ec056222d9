"""CPU tier: integration/mujoco_system_interface.patch -- the physics-thread hook and the CMake
wiring as a patch against the reference tree (B1 / N4).  Here, where /root/reference exists, the patch
must apply cleanly to a scratch copy of the three files it touches; on the GPU box (no reference)
only its content is checked."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PATCH = os.path.join(ROOT, "integration", "mujoco_system_interface.patch")
REF = "/root/reference"
FILES = ("src/mujoco_system_interface.cpp", "include/mujoco_ros2_control/mujoco_system_interface.hpp",
         "CMakeLists.txt")


def test_patch_names_the_hook_sites():
    text = open(PATCH).read()
    for f in FILES:
        assert f in text
    added = [l[1:] for l in text.splitlines() if l.startswith("+") and not l.startswith("+++")]
    joined = "\n".join(added)
    assert joined.count("scene_mirror_->refresh(mj_data_)") >= 4      # on_init, mj_forward, both mj_step sites
    assert "std::make_unique<SceneMirror>(mj_model_)" in joined
    assert "set_refresh_in_update(false)" in joined
    assert "libb2sense.so" in joined and "B2_HAVE_ROS2" in joined
    # the patch only ADDS the mirror and swaps two sources; it removes four reference lines at most
    removed = [l for l in text.splitlines() if l.startswith("-") and not l.startswith("---")]
    assert len(removed) <= 6, removed


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree is not on this machine")
def test_patch_applies_to_the_reference():
    with tempfile.TemporaryDirectory() as tmp:
        for f in FILES:
            os.makedirs(os.path.join(tmp, os.path.dirname(f)), exist_ok=True)
            shutil.copy(os.path.join(REF, f), os.path.join(tmp, f))
        r = subprocess.run(["patch", "-p1", "--forward", "-i", PATCH], cwd=tmp, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        src = open(os.path.join(tmp, FILES[0])).read()
        # both physics-loop call sites of mj_step are followed by the refresh
        assert src.count("mj_step(mj_model_, mj_data_);") == 2
        for chunk in src.split("mj_step(mj_model_, mj_data_);")[1:]:
            assert "scene_mirror_->refresh(mj_data_);" in chunk[:220]
