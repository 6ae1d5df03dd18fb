"""The drop-in boundary, executed: the reference's own programs -- main(), Update() key
handling, SDL_SaveImage -- compiled with ONLY the definition of `void Draw(screen*)` replaced by
computer-graphics_b200/host/shim/draw_{rt,rast}.inc and linked against libb200render.so
(oracle/refbuild/prog_harness.cpp, built by build_ref.sh and shipped prebuilt like oracle/_ref).
The screenshot the program saves must equal, bit for bit, the one the unmodified program saves
for the same key presses, and for the raytracer the repository's own raytracer/screenshot.bmp."""
import numpy as np
import pytest

import helpers as h
from conftest import load_golden

pytestmark = pytest.mark.gpu


def need(*names):
    for n in names:
        if not h.have_ref(n):
            pytest.skip(f"oracle/_ref/{n} not built")


def test_raytracer_program_with_dropin_draw_reproduces_screenshot_bmp():
    """raytracer/Source/skeleton.cpp:85-102: two frames, the second after one press of UP
    (cameraPos.z = -3 + 0.1f, :215) -- the pose of raytracer/screenshot.bmp (KAT #2)."""
    need("libprog_rt_dropin.so")
    shot = h.prog_run("libprog_rt_dropin.so", [[], ["UP"]])
    want = load_golden("rt_screenshot_320x256.npz")["argb"]
    assert shot.shape == want.shape
    assert np.array_equal(shot, want)


@pytest.mark.parametrize("script", [[[]], [["RIGHT", "w", "w"], ["m", "i"]], [["n", "n", "DOWN", "q"]]])
def test_raytracer_program_dropin_vs_reference_program(script):
    need("libprog_rt_dropin.so", "libprog_rt_ref.so")
    got = h.prog_run("libprog_rt_dropin.so", script)
    want = h.prog_run("libprog_rt_ref.so", script)
    assert np.array_equal(got, want), np.count_nonzero(got != want)


@pytest.mark.parametrize("size,script", [
    ("64x48", [["g"] * 95]),                                 # first frame: indirect light 0.15 on entry (:54)
    ("64x48", [["g"] * 95, ["1", "m", "UP"], ["x", "d"]]),   # later frames: 0.2 -+ 0.005, yaw, camera, light moves
    ("900x720", [[]]),                                       # BASELINE config 2, the program's very first frame
    ("900x720", [[], ["n", "LEFT", "2", "2"]]),
])
def test_rasteriser_program_dropin_vs_reference_program(size, script):
    """rasteriser/Source/skeleton.cpp:126-181 with setting = settingBoxes = 0 (untextured model)."""
    need(f"libprog_rast_dropin_{size}.so", f"libprog_rast_ref_{size}.so")
    got = h.prog_run(f"libprog_rast_dropin_{size}.so", script)
    want = h.prog_run(f"libprog_rast_ref_{size}.so", script)
    assert np.array_equal(got, want), np.count_nonzero(got != want)
