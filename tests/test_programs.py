"""CPU pins of the program harness (oracle/refbuild/prog_harness.cpp): the UNMODIFIED reference
programs, run headless through their own main() / Update() with scripted key presses."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import helpers as h
from conftest import load_golden


def test_reference_raytracer_program_reproduces_screenshot_bmp():
    """KAT #2 end to end: main() (raytracer/Source/skeleton.cpp:85-102), one press of UP handled
    by Update() (:215), the reference's Draw, SDL_SaveImage -> raytracer/screenshot.bmp."""
    if not h.have_ref("libprog_rt_ref.so"):
        pytest.skip("oracle/_ref not built")
    shot = h.prog_run("libprog_rt_ref.so", [[], ["UP"]])
    assert np.array_equal(shot, load_golden("rt_screenshot_320x256.npz")["argb"])


def test_reference_rasteriser_program_first_frame_sees_the_startup_indirect_light():
    """The program's first Draw runs with indirectLightPowerPerArea = 0.15 (:54) for its first
    shaded fragment; from the second frame on the global is 0.2 (:585).  The oracle restates it."""
    if not h.have_ref("libprog_rast_ref_64x48.so"):
        pytest.skip("oracle/_ref not built")
    W, H = 64, 48
    f = 512.0 - 5.0 * 95
    room, boxes = h.ref_rast_testmodel(W, H)
    first = h.prog_run("libprog_rast_ref_64x48.so", [["g"] * 95])
    second = h.prog_run("libprog_rast_ref_64x48.so", [["g"] * 95, []])
    o15 = h.oracle_rast_draw(W, H, f, h.DEFAULT_RAST_CAM, h.identity_R(), dict(h.DEFAULT_RAST_LIGHT, indirect=(0.15,) * 3), room, boxes)
    o20 = h.oracle_rast_draw(W, H, f, h.DEFAULT_RAST_CAM, h.identity_R(), h.DEFAULT_RAST_LIGHT, room, boxes)
    assert np.array_equal(first, o15["argb"])
    assert np.array_equal(second, o20["argb"])


def test_dropin_programs_link_the_product_library():
    """The drop-in flavour is the same translation unit with the shim's Draw, linked against
    libb200render.so (no compute here: the executed comparison is tests/test_dropin_gpu.py)."""
    for name in ("libprog_rt_dropin.so", "libprog_rast_dropin_900x720.so"):
        if not h.have_ref(name):
            pytest.skip("oracle/_ref not built")
        path = os.path.join(h.REF_DIR, name)
        out = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
        assert "libb200render.so" in out and "not found" not in out, out
        syms = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True).stdout
        assert "prog_run" in syms and "reference_Draw" in syms and "Draw" in syms
