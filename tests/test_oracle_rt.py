"""CPU tests: pin the plain-C RT oracle against the reference's golden image,
the committed outputs of the compiled reference, and (when present) the compiled
reference itself."""
import numpy as np
import pytest

import helpers as h
from conftest import load_golden


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_layouts():
    assert h.RT_TRI.itemsize == 76 and h.RT_SPHERE.itemsize == 44 and h.RAST_TRI.itemsize == 84


def test_kat2_golden_screenshot(cornell_rt):
    """KAT #2: raytracer/screenshot.bmp == Draw at cameraPos (0,0,-3+0.1f,1)
    (one UP key press, raytracer/Source/skeleton.cpp:216-217), bit-exact."""
    tris, sph = cornell_rt
    gold = load_golden("rt_screenshot_320x256.npz")["argb"]
    cam = h.f32(0, 0, np.float32(-3.0) + np.float32(0.1), 1)
    o = h.oracle_rt_render(320, 256, 256.0, cam, h.identity_R(), h.lights_array(h.DEFAULT_RT_LIGHTS), tris, sph)
    assert np.array_equal(o["argb"], gold)
    # Cornell default config ray counts quoted in SURVEY.md / BASELINE.md
    o2 = h.oracle_rt_render(320, 256, 256.0, h.f32(0, 0, -3, 1), h.identity_R(),
                            h.lights_array(h.DEFAULT_RT_LIGHTS), tris, sph, want=("index",))
    assert o2["primary"] == 737280 and o2["shadow"] == 589823
    assert np.count_nonzero(o2["index"] == -1) > 0  # the sphere is visible


@pytest.mark.parametrize("name", ["cornell_default_160x128", "cornell_yaw_96x64", "random40_96x64"])
def test_oracle_matches_committed_reference_outputs(name):
    g = load_golden(f"rt_ref_{name}.npz")
    tris = g["tris"].view(h.RT_TRI)
    sph = g["spheres"].view(h.RT_SPHERE)
    o = h.oracle_rt_render(int(g["W"]), int(g["H"]), float(g["focal"]), g["cam"], g["R"], g["lights"], tris, sph)
    assert np.array_equal(bits(o["rgb"]), bits(g["rgb"]))
    assert np.array_equal(o["argb"], g["argb"])
    assert np.array_equal(bits(o["dist"]), bits(g["dist"]))
    assert np.array_equal(o["index"], g["index"])


def test_cornell_bytes_match_reference_loader(cornell_rt):
    if not h.have_ref("libref_rt.so"):
        pytest.skip("oracle/_ref not built")
    tris, sph = h.ref_rt_testmodel()
    assert tris.tobytes() == cornell_rt[0].tobytes()
    assert sph.tobytes()[:32] == cornell_rt[1].tobytes()[:32]  # Sphere::normal is uninitialised


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_oracle_vs_compiled_reference_random(seed):
    if not h.have_ref("libref_rt.so"):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(100 + seed)
    tris, sph = h.random_rt_scene(int(rng.integers(1, 60)), seed, n_spheres=int(rng.integers(0, 3)))
    W, H = int(rng.integers(17, 80)), int(rng.integers(9, 60))
    cam = h.f32(*rng.uniform(-0.3, 0.3, 2), -2.5, 1)
    R = h.yaw_R(float(rng.uniform(-0.5, 0.5)))
    L = h.lights_array([((0.2, -0.4, -0.9, 1), (10, 12, 14)), ((-0.5, 0.3, -1.2, 1), (3, 2, 1))][: 1 + seed % 2])
    o = h.oracle_rt_render(W, H, 70.0, cam, R, L, tris, sph if len(sph) else None)
    r = h.ref_rt_draw(W, H, 70.0, cam, R, L, tris, sph if len(sph) else np.zeros(0, h.RT_SPHERE))
    t = h.ref_rt_trace(W, H, 70.0, cam, R, tris, sph if len(sph) else np.zeros(0, h.RT_SPHERE))
    assert np.array_equal(bits(o["rgb"]), bits(r["rgb"]))
    assert np.array_equal(bits(o["dist"]), bits(t["dist"]))
    assert np.array_equal(o["index"], t["index"])


def test_oracle_bands_tile_the_frame(cornell_rt):
    tris, sph = cornell_rt
    L = h.lights_array(h.DEFAULT_RT_LIGHTS)
    full = h.oracle_rt_render(64, 48, 48.0, h.f32(0, 0, -3, 1), h.identity_R(), L, tris, sph)
    a = h.oracle_rt_render(64, 48, 48.0, h.f32(0, 0, -3, 1), h.identity_R(), L, tris, sph, 0, 17)
    b = h.oracle_rt_render(64, 48, 48.0, h.f32(0, 0, -3, 1), h.identity_R(), L, tris, sph, 17, 48)
    assert np.array_equal(full["rgb"][:17], a["rgb"][:17]) and np.array_equal(full["rgb"][17:], b["rgb"][17:])
    assert full["shadow"] == a["shadow"] + b["shadow"]
