"""GPU parity tests for the raytracer path: the CUDA kernels (through the C ABI)
against the plain-C oracle on the same inputs, and against the committed golden
fixtures.  Bar: bit-exact float colour, distance and index."""
import numpy as np
import pytest

import helpers as h
from conftest import load_golden

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def check_equal(got, want, what):
    assert np.array_equal(got["index"], want["index"]), f"{what}: index differs at {np.count_nonzero(got['index'] != want['index'])} px"
    assert np.array_equal(bits(got["depth"]), bits(want["dist"])), f"{what}: distance differs at {np.count_nonzero(bits(got['depth']) != bits(want['dist']))} px"
    nb = np.count_nonzero((bits(got["rgb"]) != bits(want["rgb"])).any(axis=-1))
    assert nb == 0, f"{what}: colour differs at {nb} px"


def run_both(b200, renderer, tris, sph, W, H, focal, cam, R, lights, what, brute=True):
    c = b200.make_camera(cam, focal, R, W, H)
    want = h.oracle_rt_render(W, H, focal, cam, R, h.lights_array(lights), tris, sph if sph is not None and len(sph) else None)
    renderer.set_option(b200.OPT_RT_BRUTEFORCE, 0)
    renderer.set_option(b200.OPT_RT_GRID, 2)          # the scene-streaming kernel, whatever the scene size
    try:
        got = renderer.render_raytrace(tris, sph, c, lights)
        st = renderer.stats()
    finally:
        renderer.set_option(b200.OPT_RT_GRID, 0)
    check_equal(got, want, what + " [filtered]")
    assert st["primary_rays"] == want["primary"] and st["shadow_rays"] == want["shadow"]
    # the same frame through the direction grids (automatic only for large scenes)
    # (twice: from the second gridded frame of a scene size on, the lists are filled from the (cell, record) pairs
    # the counting pass wrote down instead of by a second descent)
    renderer.set_option(b200.OPT_RT_GRID, 1)
    try:
        for frame in range(2):
            got_g = renderer.render_raytrace(tris, sph, c, lights)
            st_g = renderer.stats()
            check_equal(got_g, want, what + f" [filtered, grids, frame {frame}]")
            assert st_g["shadow_rays"] == want["shadow"]
    finally:
        renderer.set_option(b200.OPT_RT_GRID, 0)
    if brute:
        renderer.set_option(b200.OPT_RT_BRUTEFORCE, 1)
        got2 = renderer.render_raytrace(tris, sph, c, lights)
        st2 = renderer.stats()
        renderer.set_option(b200.OPT_RT_BRUTEFORCE, 0)
        check_equal(got2, want, what + " [bruteforce]")
        assert st2["shadow_rays"] == want["shadow"]
    return got, want, st


def test_kat2_golden_screenshot(b200, renderer, cornell_rt):
    """The reference's own golden image, bit-exact through draw_raytrace."""
    tris, sph = cornell_rt
    gold = load_golden("rt_screenshot_320x256.npz")["argb"]
    cam = h.f32(0, 0, np.float32(-3.0) + np.float32(0.1), 1)
    c = b200.make_camera(cam, 256.0, h.identity_R(), 320, 256)
    argb = renderer.draw_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS)
    assert np.array_equal(argb, gold)


def test_cornell_default_config(b200, renderer, cornell_rt):
    """BASELINE config 1: 320x256, f=256, cam (0,0,-3,1)."""
    tris, sph = cornell_rt
    _, want, st = run_both(b200, renderer, tris, sph, 320, 256, 256.0, h.f32(0, 0, -3, 1), h.identity_R(),
                           h.DEFAULT_RT_LIGHTS, "cornell 320x256")
    assert st["primary_rays"] == 737280 and st["shadow_rays"] == 589823
    # most pairs are decided by the filter, not by the reference arithmetic
    assert st["exact_evals"] < 0.2 * st["prim_tests"]


@pytest.mark.parametrize("name", ["cornell_default_160x128", "cornell_yaw_96x64", "random40_96x64"])
def test_committed_reference_outputs(b200, renderer, name):
    g = load_golden(f"rt_ref_{name}.npz")
    tris = g["tris"].view(h.RT_TRI).copy()
    sph = g["spheres"].view(h.RT_SPHERE).copy()
    W, H = int(g["W"]), int(g["H"])
    c = b200.make_camera(g["cam"], float(g["focal"]), g["R"], W, H)
    L = [(g["lights"][:4], g["lights"][4:7])]
    got = renderer.render_raytrace(tris, sph, c, L)
    check_equal(got, dict(rgb=g["rgb"], dist=g["dist"], index=g["index"]), name)
    assert np.array_equal(renderer.draw_raytrace(tris, sph, c, L), g["argb"])


@pytest.mark.parametrize("seed", range(6))
def test_random_scenes(b200, renderer, seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 700))
    tris, sph = h.random_rt_scene(n, seed, size=float(rng.uniform(0.05, 0.8)), n_spheres=int(rng.integers(0, 3)))
    W, H = int(rng.integers(17, 90)), int(rng.integers(9, 70))
    cam = h.f32(*rng.uniform(-0.3, 0.3, 2), -2.5, 1)
    R = h.yaw_R(float(rng.uniform(-0.5, 0.5)))
    lights = [((0.2, -0.4, -0.9, 1), (10, 12, 14)), ((-0.5, 0.3, -1.2, 1), (3, 2, 1)),
              ((0.0, 0.9, -0.2, 1), (5, 5, 5))][: 1 + seed % 3]
    run_both(b200, renderer, tris, sph, W, H, 70.0, cam, R, lights, f"random seed {seed}")


def test_camera_inside_scene_and_yaw(b200, renderer, cornell_rt):
    tris, sph = cornell_rt
    run_both(b200, renderer, tris, sph, 128, 96, 90.0, h.f32(0.3, 0.2, -0.5, 1), h.yaw_R(0.7),
             h.DEFAULT_RT_LIGHTS, "inside+yaw")


def test_light_in_a_triangle_plane(b200, renderer, cornell_rt):
    """Degenerate origin: the light lies exactly in the ceiling's plane (y = -1)."""
    tris, sph = cornell_rt
    run_both(b200, renderer, tris, sph, 96, 80, 80.0, h.f32(0, 0, -3, 1), h.identity_R(),
             [((0.0, -1.0, -0.3, 1), (14, 14, 14))], "light in plane")


def test_edge_cases(b200, renderer, cornell_rt):
    tris, sph = cornell_rt
    L = h.DEFAULT_RT_LIGHTS
    # no triangles, only the sphere; no spheres; nothing at all; no lights
    run_both(b200, renderer, np.zeros(0, h.RT_TRI), sph, 40, 30, 30.0, h.f32(0, 0, -3, 1), h.identity_R(), L, "sphere only")
    run_both(b200, renderer, tris, np.zeros(0, h.RT_SPHERE), 40, 30, 30.0, h.f32(0, 0, -3, 1), h.identity_R(), L, "no spheres")
    got, _, _ = run_both(b200, renderer, np.zeros(0, h.RT_TRI), np.zeros(0, h.RT_SPHERE), 19, 7, 30.0,
                         h.f32(0, 0, -3, 1), h.identity_R(), L, "empty scene")
    assert not got["rgb"].any() and (got["index"] == h.INDEX_MISS).all() and np.isinf(got["depth"]).all()
    run_both(b200, renderer, tris, sph, 33, 17, 30.0, h.f32(0, 0, -3, 1), h.identity_R(), [], "no lights")
    # a single pixel, and a 1-row frame
    run_both(b200, renderer, tris, sph, 1, 1, 1.0, h.f32(0, 0, -3, 1), h.identity_R(), L, "1x1")
    run_both(b200, renderer, tris, sph, 257, 1, 128.0, h.f32(0, 0, -3, 1), h.identity_R(), L, "257x1")


def test_bands_tile_the_frame(b200, renderer, cornell_rt):
    tris, sph = cornell_rt
    cam = h.f32(0, 0, -3, 1)
    c = b200.make_camera(cam, 100.0, h.identity_R(), 120, 100)
    full = renderer.render_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS)
    n_shadow = renderer.stats()["shadow_rays"]
    parts, acc = [], 0
    for r0, r1 in [(0, 13), (13, 50), (50, 51), (51, 100)]:
        parts.append(renderer.render_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS, r0, r1))
        acc += renderer.stats()["shadow_rays"]
    for key in ("rgb", "depth", "index"):
        assert np.array_equal(np.concatenate([p[key] for p in parts]), full[key])
    assert acc == n_shadow


def test_many_triangles_multi_tile(b200, renderer):
    """More triangles than one shared-memory tile (256): exercises the TMA ring."""
    tris, sph = h.random_rt_scene(1500, 42, size=0.15, n_spheres=1)
    run_both(b200, renderer, tris, sph, 64, 48, 60.0, h.f32(0, 0, -2.6, 1), h.identity_R(),
             h.DEFAULT_RT_LIGHTS, "1500 tris")


def test_invalid_arguments(b200, renderer, cornell_rt):
    tris, sph = cornell_rt
    c = b200.make_camera(h.f32(0, 0, -3, 1), 10.0, h.identity_R(), 0, 10)
    with pytest.raises(b200.B200Error):
        renderer.render_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS)
    c = b200.make_camera(h.f32(0, 0, -3, 1), 10.0, h.identity_R(), 16, 16)
    with pytest.raises(b200.B200Error):
        renderer.render_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS, 5, 40)
    with pytest.raises(b200.B200Error):
        renderer.render_raytrace(tris, sph, c, [((0, 0, 0, 1), (1, 1, 1))] * 9)


def test_grid_tessellated_scene(b200, renderer):
    """BASELINE config 5's generator at a size the oracle finishes in seconds: large enough
    for the direction grids to switch on by themselves."""
    tris, sph = b200.scene_cornell_rt_tessellated(10)        # 2800 triangles
    assert len(tris) == 2800
    run_both(b200, renderer, tris, sph, 150, 110, 75.0, h.f32(0, 0, -3, 1), h.identity_R(),
             h.DEFAULT_RT_LIGHTS, "tessellated 2800", brute=False)
    # automatic mode: with the lists the kernel launches next to the binning kernels; without them only prep + render
    c = b200.make_camera(h.f32(0, 0, -3, 1), 75.0, h.identity_R(), 150, 110)
    renderer.render_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS, want=("rgb",))
    assert renderer.stats()["kernel_launches"] >= 7
    run_both(b200, renderer, tris, sph, 97, 61, 50.0, h.f32(0.2, -0.1, -0.7, 1), h.yaw_R(-0.6),
             [((0.3, -0.6, -0.2, 1), (9, 9, 9)), ((-0.4, 0.5, -0.9, 1), (4, 5, 6))], "tessellated, inside, 2 lights",
             brute=False)


def test_grid_distance_ties_go_to_the_lowest_index(b200, renderer, cornell_rt):
    """Coincident copies of every triangle with different colours: equal distances, and the
    reference keeps the first (strict `<`, skeleton.cpp:313).  The grid lists are unordered."""
    tris, sph = cornell_rt
    dup = np.concatenate([tris, tris, tris]).copy()
    dup["color"][len(tris):2 * len(tris)] = (0.9, 0.1, 0.9)
    dup["color"][2 * len(tris):] = (0.1, 0.9, 0.1)
    got, want, _ = run_both(b200, renderer, dup, sph, 90, 70, 60.0, h.f32(0, 0, -3, 1), h.identity_R(),
                            h.DEFAULT_RT_LIGHTS, "coincident triangles")
    assert got["index"].max() < len(tris)


def test_grid_full_size_matches_streaming_kernel(b200, renderer):
    """BASELINE config 5 geometry (100 800 triangles) on a 480x270 frame: the grid kernel
    against the scene-streaming kernel (itself checked against the oracle above)."""
    tris, sph = b200.scene_cornell_rt_tessellated(60)
    c = b200.make_camera(h.f32(0, 0, -3, 1), 192.0, h.identity_R(), 480, 270)
    renderer.set_option(b200.OPT_RT_GRID, 2)
    try:
        a = renderer.render_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS)
        sa = renderer.stats()
    finally:
        renderer.set_option(b200.OPT_RT_GRID, 0)
    b = renderer.render_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS)
    sb = renderer.stats()
    for k in ("rgb", "depth", "index"):
        assert np.array_equal(a[k], b[k]), k
    assert sa["shadow_rays"] == sb["shadow_rays"] and sb["kernel_launches"] > sa["kernel_launches"]


def test_sliced_framebuffer_return(b200, renderer, cornell_rt):
    """Frames of three megapixels or more come back from draw_raytrace(_band) in slices that overlap
    the rendering; the packed frame must equal the quantised float frame of the plain path."""
    tris, sph = cornell_rt
    W, H = 2064, 1700       # 3.5 Mpixel (slices from 3 Mi pixels; the band below: 3.02 Mi), not a multiple of the 16-row blocks
    c = b200.make_camera(h.f32(0, 0, -3, 1), 1050.0, h.identity_R(), W, H)
    want = b200.quantise(renderer.render_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS, want=("rgb",))["rgb"])
    got = renderer.draw_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS)
    assert np.array_equal(got, want)
    band = np.zeros((H - 37 - 100, W), np.uint32)
    renderer.draw_raytrace_band(tris, sph, c, h.DEFAULT_RT_LIGHTS, 37, H - 100, band.ctypes.data)
    assert np.array_equal(band, want[37:H - 100])


@pytest.mark.parametrize("mode", ["filtered", "grids", "bruteforce"])
def test_interleaved_blocks_tile_the_frame(b200, renderer, cornell_rt, mode):
    """B200_OPT_RT_INTERLEAVE_N/R: n contexts (here one, three times) render the 16-row blocks
    b % n == r of the same row range into one full-frame buffer; together they give the frame."""
    import torch
    tris, sph = cornell_rt
    W, H = 150, 117                    # 8 blocks of 16 rows, the last one partial
    c = b200.make_camera(h.f32(0, 0, -3, 1), 100.0, h.identity_R(), W, H)
    renderer.set_option(b200.OPT_RT_BRUTEFORCE, 1 if mode == "bruteforce" else 0)
    renderer.set_option(b200.OPT_RT_GRID, 1 if mode == "grids" else 2)
    try:
        full = renderer.render_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS)
        rays = renderer.stats()
        renderer.rt_upload_scene(tris, sph)
        for row0, row1 in ((0, H), (5, 101)):
            rgb = torch.full((H, W, 3), float("nan"), device="cuda")
            depth = torch.full((H, W), float("nan"), device="cuda")
            index = torch.full((H, W), 12345, dtype=torch.int32, device="cuda")
            primary = shadow = 0
            for r in range(3):
                renderer.set_option(b200.OPT_RT_INTERLEAVE_N, 3)
                renderer.set_option(b200.OPT_RT_INTERLEAVE_R, r)
                renderer.rt_render_device(c, h.DEFAULT_RT_LIGHTS, row0, row1, rgb.data_ptr(), depth.data_ptr(), index.data_ptr())
                st = renderer.stats()
                primary += st["primary_rays"]; shadow += st["shadow_rays"]
            assert np.array_equal(bits(rgb.cpu().numpy()[row0:row1]), bits(full["rgb"][row0:row1]))
            assert np.array_equal(bits(depth.cpu().numpy()[row0:row1]), bits(full["depth"][row0:row1]))
            assert np.array_equal(index.cpu().numpy()[row0:row1], full["index"][row0:row1])
            assert (index.cpu().numpy()[:row0] == 12345).all() and (index.cpu().numpy()[row1:] == 12345).all()
            assert primary == W * (row1 - row0) * 9
            if (row0, row1) == (0, H):
                assert shadow == rays["shadow_rays"]
    finally:
        renderer.set_option(b200.OPT_RT_INTERLEAVE_N, 1)
        renderer.set_option(b200.OPT_RT_BRUTEFORCE, 0)
        renderer.set_option(b200.OPT_RT_GRID, 0)


def test_unnormalised_normals(b200, renderer):
    """The shadow ray starts at hit + 1e-5 * normal (:394) with whatever normal the caller
    supplies; the filter margins scale with the largest normal component."""
    tris, sph = h.random_rt_scene(120, 77, size=0.5, n_spheres=1)
    tris["normal"][:, :3] *= np.float32(43.0)
    tris["normal"][::7, :3] *= np.float32(0.01)
    run_both(b200, renderer, tris, sph, 80, 60, 60.0, h.f32(0.1, -0.1, -2.5, 1), h.yaw_R(0.2),
             h.DEFAULT_RT_LIGHTS, "unnormalised normals")


def test_tessellated_frame_matches_the_plain_cornell_frame(b200, renderer, cornell_rt):
    """the tessellated box is the same surface, so its frame must equal the
    plain Cornell frame except where a sample falls on a tessellation edge (ties / ulp-level
    differences of the hit distance): a built-in cross-check of the large-scene path."""
    tris, sph = cornell_rt
    tess, _ = b200.scene_cornell_rt_tessellated(12)       # 4032 triangles: direction grids on
    # (not the default pose: with cam z = -3 and an integer focal length the outline of the box
    # falls exactly on pixel-centre rays and every outline sample is a tie)
    c = b200.make_camera(h.f32(0.013, -0.021, -3.02, 1), 163.7, h.identity_R(), 200, 160)
    a = renderer.render_raytrace(tris, sph, c, h.DEFAULT_RT_LIGHTS)
    b = renderer.render_raytrace(tess, sph, c, h.DEFAULT_RT_LIGHTS)
    miss_a, miss_b = np.isinf(a["depth"]), np.isinf(b["depth"])
    assert np.count_nonzero(miss_a != miss_b) <= 0.002 * miss_a.size            # same silhouette up to edge samples
    hit = ~miss_a & ~miss_b
    assert np.allclose(a["depth"][hit], b["depth"][hit], rtol=2e-5, atol=0)
    # spheres keep their index; triangle hits map back to their parent (144 children each)
    ia, ib = a["index"][hit], b["index"][hit]
    parent = np.where(ib >= 0, ib // 144, ib)
    assert np.count_nonzero(parent != ia) <= 0.002 * hit.sum()                  # ties on shared parent edges
    diff = np.abs(a["rgb"] - b["rgb"]).max(axis=-1)
    assert np.count_nonzero(diff > 1e-4) <= 0.01 * diff.size                    # shadow / tessellation edges only
    assert np.median(diff) <= 1e-6


def test_planned_frames_split_blocks(b200, renderer):
    """Gridded frames that follow a frame of the same shape run from a plan made of that frame's block costs:
    expensive blocks first, split into eight launches of one pixel column per 8x4 patch.  With the split
    threshold at the mean (b200_debug_rt_plan_heavy) about half the blocks of a small frame are split; every
    frame must still be the oracle's, and the ray counts must not change."""
    tris, sph = b200.scene_cornell_rt_tessellated(8)
    W, H, f = 150, 110, 75.0
    cam, R = h.f32(0.1, 0.0, -2.2, 1), h.yaw_R(0.3)
    lights = [((0.3, -0.6, -0.2, 1), (9, 9, 9)), ((-0.4, 0.5, -0.9, 1), (4, 5, 6))]
    want = h.oracle_rt_render(W, H, f, cam, R, h.lights_array(lights), tris, sph)
    c = b200.make_camera(cam, f, R, W, H)
    lib = b200.load_library()
    renderer.set_option(b200.OPT_RT_GRID, 1)
    try:
        for n_l in (1, 2):
            want_l = want if n_l == 2 else h.oracle_rt_render(W, H, f, cam, R, h.lights_array(lights[:1]), tris, sph)
            for heavy in (1, 2, 0):
                assert lib.b200_debug_rt_plan_heavy(renderer.ctx, heavy) == 0
                launches = []
                for frame in range(3):
                    got = renderer.render_raytrace(tris, sph, c, lights[:n_l])
                    st = renderer.stats()
                    check_equal(got, want_l, f"{n_l} light(s), split above {heavy} x mean, frame {frame}")
                    assert st["primary_rays"] == want_l["primary"] and st["shadow_rays"] == want_l["shadow"]
                    launches.append(st["kernel_launches"])
            renderer.set_option(b200.OPT_RT_PLAN, 0)
            got = renderer.render_raytrace(tris, sph, c, lights[:n_l])
            check_equal(got, want_l, "unplanned")
            renderer.set_option(b200.OPT_RT_PLAN, 1)
    finally:
        lib.b200_debug_rt_plan_heavy(renderer.ctx, 0)
        renderer.set_option(b200.OPT_RT_PLAN, 1)
        renderer.set_option(b200.OPT_RT_GRID, 0)
