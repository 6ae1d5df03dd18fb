"""GPU parity tests of the rasteriser's texture branches (rasteriser/Source/skeleton.cpp:588-645,
findU / findV :1756-1825) through the C ABI (rast_set_textures), against the plain-C oracle -- which
tests/test_oracle_rast.py pins bit for bit on the unmodified reference -- on synthetic images.  Bar:
bit-exact depth (holes included), owner, colour buffers, shadow mask and final colour."""
import numpy as np
import pytest

import helpers as h

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def tex():
    return h.synthetic_textures()


@pytest.fixture()
def textured(renderer, tex):
    renderer.set_textures(tex)
    yield renderer
    renderer.set_textures(None)
    h.oracle_rast_set_textures(None)


def cornell(setting, setting_boxes):
    room, boxes = h.golden_cornell_rast()
    room["texture"] = setting
    boxes["texture"] = setting_boxes
    return room, boxes


def check_draw(b200, r, tex, W, H, f, cam_pos, yaw, light, room, boxes, what, buffers=True):
    R = h.yaw_R(yaw) if yaw != 0 else h.identity_R()
    h.oracle_rast_set_textures(tex, cam_pos, R, yaw)
    want = h.oracle_rast_draw(W, H, f, cam_pos, R, light, room, boxes)
    cam = b200.make_camera(cam_pos, f, R, W, H)
    L = b200.make_rast_light(light["pos"], light["power"], light["indirect"])
    r.set_option(b200.OPT_RAST_PATH, 0)
    got = r.render_raster(room, boxes, cam, L)
    assert np.array_equal(got["index"], want["index"]), f"{what}: owner differs at {np.count_nonzero(got['index'] != want['index'])} px"
    assert np.array_equal(bits(got["depth"]), bits(want["depth"])), f"{what}: depth"
    assert np.array_equal(bits(got["rgb"]), bits(want["rgb"])), f"{what}: final colour at {np.count_nonzero(bits(got['rgb']) != bits(want['rgb']))} words"
    assert r.stats()["fragments"] == want["fragments"]
    if buffers:   # the path that keeps the reference's six buffers
        r.set_option(b200.OPT_RAST_PATH, 1)
        for ts in (5, 3):
            r.set_option(b200.OPT_RAST_TILE_LOG2, ts)
            got = r.render_raster(room, boxes, cam, L)
            buf = r.raster_read_buffers(W, H)
            assert np.array_equal(bits(got["depth"]), bits(want["depth"])), f"{what} [tile {1 << ts}]: depth"
            assert np.array_equal(buf["shadow"], want["shadow"]), f"{what} [tile {1 << ts}]: shadow mask"
            for key in ("screen", "low", "high"):
                assert np.array_equal(bits(buf[key]), bits(want[key])), f"{what} [tile {1 << ts}]: {key}"
            assert np.array_equal(bits(got["rgb"]), bits(want["rgb"])), f"{what} [tile {1 << ts}]: final colour"
        r.set_option(b200.OPT_RAST_TILE_LOG2, 5)
        r.set_option(b200.OPT_RAST_PATH, 0)
    return want


@pytest.mark.parametrize("setting,setting_boxes", [(2, 1), (1, 3), (3, 2), (0, 1), (3, 3)])
@pytest.mark.parametrize("cam_pos,yaw", [((0.0, 0.0, -3.001, 1.0), 0.0), ((0.1, -0.05, -2.6, 1.0), 0.174533),
                                         ((0.0, 0.0, -3.001, 1.0), -0.349066)])
def test_textured_cornell_box(b200, textured, tex, setting, setting_boxes, cam_pos, yaw):
    """The reference's own scene with its texture settings (it ships with 2, 1: metal grill on the room,
    marble on the boxes), straight on and through the `yaw != 0` branch of findU / findV."""
    room, boxes = cornell(setting, setting_boxes)
    want = check_draw(b200, textured, tex, 320, 240, 120.0, cam_pos, yaw, h.DEFAULT_RAST_LIGHT, room, boxes,
                      f"cornell tex {setting}/{setting_boxes} yaw {yaw}")
    if setting_boxes in (2, 3):   # holes that let an earlier colour show through a cleared depth
        assert np.count_nonzero((want["depth"] == 0) & (want["index"] >= 0)) > 100


def test_textured_cornell_900x720_default_settings(b200, textured, tex):
    """BASELINE config 2's resolution with the settings the reference ships with."""
    room, boxes = cornell(2, 1)
    check_draw(b200, textured, tex, 900, 720, 512.0, h.DEFAULT_RAST_CAM, 0.0, h.DEFAULT_RAST_LIGHT, room, boxes,
               "cornell 900x720 tex 2/1", buffers=False)


def test_entry_indirect_light_skips_holes(b200, textured, tex):
    """indirectLightPowerPerArea's entry value reaches the first SHADED fragment (:585): a hole is
    accepted but not shaded (:603, :625), so the search goes on behind it."""
    room, boxes = cornell(2, 3)
    light = dict(h.DEFAULT_RAST_LIGHT, indirect=(0.15, 0.15, 0.15))
    check_draw(b200, textured, tex, 320, 240, 120.0, h.DEFAULT_RAST_CAM, 0.0, light, room, boxes, "entry 0.15, tex 2/3")
    light = dict(h.DEFAULT_RAST_LIGHT, indirect=(0.21, 0.2, 0.19))
    check_draw(b200, textured, tex, 64, 48, 24.0, (0.05, 0.0, -2.8, 1.0), 0.174533, light, room, boxes, "entry mixed, tex 2/3")


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_textured_random_clipped_lists(b200, textured, tex, seed):
    """Tier 1 on random camera-space lists: every texture on every object index, shadow triangles between
    them, triangles far outside the unit box (findU / findV wrap instead of reading out of bounds)."""
    W, H, f = 256, 192, 150.0
    rng = np.random.default_rng(100 + seed)
    clipped = h.random_clipped_list(400, seed, W, H, f, shadow_frac=0.15, size=0.6)
    clipped["texture"] = rng.integers(0, 4, len(clipped))
    clipped["index"] = rng.integers(0, 6, len(clipped))     # 5: no branch of findU / findV, texel (0, 0)
    clipped["index"][::7] = 0x01000003                      # garbage (the reference leaves one index uninitialised,
    clipped["index"][3::11] = -7                            # TestModelH.h:254-257): no branch either
    clipped["texture"][clipped["color"][:, 0] < 0] = 0
    cam_pos, yaw = (0.2, -0.1, -1.5, 1.0), (0.0 if seed == 1 else 0.3)
    R = h.yaw_R(yaw) if yaw != 0 else h.identity_R()
    light_cam = h.f32(0.1, -0.4, 1.2, 1.0)
    h.oracle_rast_set_textures(tex, cam_pos, R, yaw)
    want = h.oracle_rast_draw_clipped(W, H, f, light_cam, h.DEFAULT_RAST_LIGHT, clipped)
    cam = b200.make_camera(cam_pos, f, R, W, H)
    L = b200.make_rast_light(light_cam, h.DEFAULT_RAST_LIGHT["power"], h.DEFAULT_RAST_LIGHT["indirect"])
    for path in (0, 1):
        textured.set_option(b200.OPT_RAST_PATH, path)
        got = textured.render_raster_clipped(clipped, cam, L)
        assert np.array_equal(got["index"], want["index"]), f"path {path}: owner"
        assert np.array_equal(bits(got["depth"]), bits(want["depth"])), f"path {path}: depth"
        assert np.array_equal(bits(got["rgb"]), bits(want["rgb"])), f"path {path}: final colour"
    textured.set_option(b200.OPT_RAST_PATH, 0)


def test_textured_bands_tile_the_frame(b200, textured, tex):
    """Row bands (what the devices of a multi-GPU context render) of a textured frame."""
    W, H, f = 320, 240, 120.0
    room, boxes = cornell(3, 1)
    cam_pos = h.DEFAULT_RAST_CAM
    h.oracle_rast_set_textures(tex, cam_pos, h.identity_R(), 0.0)
    want = h.oracle_rast_draw(W, H, f, cam_pos, h.identity_R(), h.DEFAULT_RAST_LIGHT, room, boxes)
    cam = b200.make_camera(cam_pos, f, h.identity_R(), W, H)
    L = b200.make_rast_light(h.DEFAULT_RAST_LIGHT["pos"], h.DEFAULT_RAST_LIGHT["power"], h.DEFAULT_RAST_LIGHT["indirect"])
    for a, b in ((0, 77), (77, 160), (160, 240)):
        got = textured.render_raster(room, boxes, cam, L, row_begin=a, row_end=b)
        assert np.array_equal(bits(got["rgb"]), bits(want["rgb"][a:b])), (a, b)
        assert np.array_equal(bits(got["depth"]), bits(want["depth"][a:b])), (a, b)
        assert np.array_equal(got["index"], want["index"][a:b]), (a, b)


def test_texture_fields_need_textures(b200, renderer, tex):
    room, boxes = cornell(2, 1)
    cam = b200.make_camera(h.DEFAULT_RAST_CAM, 24.0, h.identity_R(), 64, 48)
    L = b200.make_rast_light(h.DEFAULT_RAST_LIGHT["pos"], h.DEFAULT_RAST_LIGHT["power"], h.DEFAULT_RAST_LIGHT["indirect"])
    renderer.set_textures(None)
    with pytest.raises(b200.B200Error):
        renderer.render_raster(room, boxes, cam, L)
    renderer.set_textures(tex)
    try:
        renderer.set_option(b200.OPT_RAST_PATH, 2)          # the scatter path cannot draw holes
        with pytest.raises(b200.B200Error):
            renderer.render_raster(room[:4], boxes[:0], cam, L)
        renderer.set_option(b200.OPT_RAST_PATH, 0)
        bad = room.copy()
        bad["texture"] = 7
        with pytest.raises(b200.B200Error):
            renderer.render_raster(bad, boxes, cam, L)
        small = dict(tex, grill=tex["grill"][:512])
        with pytest.raises(b200.B200Error):
            renderer.set_textures(small)
    finally:
        renderer.set_option(b200.OPT_RAST_PATH, 0)
        renderer.set_textures(None)
    # untextured frames are unaffected afterwards
    room0, boxes0 = cornell(0, 0)
    got = renderer.render_raster(room0, boxes0, cam, L)
    want = h.oracle_rast_draw(64, 48, 24.0, h.DEFAULT_RAST_CAM, h.identity_R(), h.DEFAULT_RAST_LIGHT, room0, boxes0)
    assert np.array_equal(bits(got["rgb"]), bits(want["rgb"]))


@pytest.mark.parametrize("size,script", [
    ("64x48", [["g"] * 95]),
    ("64x48", [["g"] * 95, ["1", "m", "UP"], ["x", "d"]]),
    ("900x720", [[], ["n", "LEFT", "2"]]),
])
def test_rasteriser_program_with_the_reference_default_textures(size, script):
    """The executed drop-in on the model the reference ships with (setting = 2, settingBoxes = 1,
    TestModelH.h:9-10): main() loads its images (cv_stub.hpp's procedural ones), thresholds the opacity
    maps and draws normalMap_marble from rand(); the shim hands that state to rast_set_textures."""
    for n in (f"libprog_rast_dropin_{size}.so", f"libprog_rast_ref_{size}.so"):
        if not h.have_ref(n):
            pytest.skip(f"oracle/_ref/{n} not built")
    got = h.prog_run(f"libprog_rast_dropin_{size}.so", script, model=(2, 1))
    want = h.prog_run(f"libprog_rast_ref_{size}.so", script, model=(2, 1))
    assert np.array_equal(got, want), np.count_nonzero(got != want)
    plain = h.prog_run(f"libprog_rast_ref_{size}.so", script)
    assert not np.array_equal(plain, want)     # the textures do show


# ---- colour modes 1 / 2 (randColourSelect, :647-662; B200_OPT_RAST_COLOUR_MODE) ----
@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("W,H,f,cam_pos,yaw,indirect", [(320, 240, 120.0, h.DEFAULT_RAST_CAM, 0.0, 0.2),
                                                        (256, 192, 90.0, (0.1, -0.05, -2.6, 1.0), 0.174533, 0.15),
                                                        (900, 720, 512.0, h.DEFAULT_RAST_CAM, 0.0, 0.2)])
def test_colour_modes(b200, renderer, mode, W, H, f, cam_pos, yaw, indirect):
    """Every accepted fragment draws three rand() values in the reference's serial fragment order: the
    library numbers the accepted fragments on the device and draws from the process's rand() itself.
    Same seed: same pixels as the oracle (pinned on the unmodified reference), and the C library's stream
    ends at the same place."""
    R = h.yaw_R(yaw) if yaw else h.identity_R()
    room, boxes = cornell(0, 0)
    light = dict(h.DEFAULT_RAST_LIGHT, indirect=(indirect,) * 3)
    cam = b200.make_camera(cam_pos, f, R, W, H)
    L = b200.make_rast_light(light["pos"], light["power"], light["indirect"])
    try:
        h.oracle().oracle_rast_set_colour_mode(mode)
        h.srand(4242)
        want = h.oracle_rast_draw(W, H, f, cam_pos, R, light, room, boxes)
        next_want = h.rand()
        renderer.set_option(b200.OPT_RAST_COLOUR_MODE, mode)
        h.srand(4242)
        got = renderer.render_raster(room, boxes, cam, L)
        next_got = h.rand()
        with pytest.raises(b200.B200Error):      # the ordinals are those of the whole frame
            renderer.render_raster(room, boxes, cam, L, row_begin=8, row_end=H)
    finally:
        h.oracle().oracle_rast_set_colour_mode(0)
        renderer.set_option(b200.OPT_RAST_COLOUR_MODE, 0)
    assert np.array_equal(got["index"], want["index"])
    assert np.array_equal(bits(got["depth"]), bits(want["depth"]))
    assert np.array_equal(bits(got["rgb"]), bits(want["rgb"])), np.count_nonzero(bits(got["rgb"]) != bits(want["rgb"]))
    assert next_got == next_want
    # and the default mode is back
    got0 = renderer.render_raster(room, boxes, cam, L)
    want0 = h.oracle_rast_draw(W, H, f, cam_pos, R, light, room, boxes)
    assert np.array_equal(bits(got0["rgb"]), bits(want0["rgb"]))


def test_colour_modes_ignore_textures_and_clipped_lists(b200, textured, tex):
    """Tier 1, a random list with shadow triangles; texture fields set and textures loaded: :575 switches on
    randColourSelect before anything looks at them."""
    W, H, f = 256, 192, 150.0
    clipped = h.random_clipped_list(300, 9, W, H, f, shadow_frac=0.2, size=0.6)
    clipped["texture"] = np.where(clipped["color"][:, 0] < 0, 0, 2)
    light_cam = h.f32(0.1, -0.4, 1.2, 1.0)
    cam = b200.make_camera((0, 0, 0, 1), f, h.identity_R(), W, H)
    L = b200.make_rast_light(light_cam, h.DEFAULT_RAST_LIGHT["power"], h.DEFAULT_RAST_LIGHT["indirect"])
    try:
        h.oracle_rast_set_textures(tex, (0, 0, 0, 1), h.identity_R(), 0.0)
        h.oracle().oracle_rast_set_colour_mode(1)
        h.srand(7)
        want = h.oracle_rast_draw_clipped(W, H, f, light_cam, h.DEFAULT_RAST_LIGHT, clipped)
        textured.set_option(b200.OPT_RAST_COLOUR_MODE, 1)
        h.srand(7)
        got = textured.render_raster_clipped(clipped, cam, L)
    finally:
        h.oracle().oracle_rast_set_colour_mode(0)
        textured.set_option(b200.OPT_RAST_COLOUR_MODE, 0)
    assert np.array_equal(got["index"], want["index"])
    assert np.array_equal(bits(got["depth"]), bits(want["depth"]))
    assert np.array_equal(bits(got["rgb"]), bits(want["rgb"]))


def test_rasteriser_program_space_key_cycles_the_colour_modes():
    """SPACE (:405-408) cycles randColourSelect; frames in modes 1 / 2 draw from rand(), whose stream also fed
    normalMap_marble at start-up (:157-168, srand(1) in the harness): the drop-in program stays in step with
    the unmodified one frame after frame."""
    size = "64x48"
    for n in (f"libprog_rast_dropin_{size}.so", f"libprog_rast_ref_{size}.so"):
        if not h.have_ref(n):
            pytest.skip(f"oracle/_ref/{n} not built")
    for script in ([["g"] * 95 + [" "]], [["g"] * 95, [" "], ["m"], [" ", "UP"]], [["g"] * 95, [" ", " "], [" "], ["1"]]):
        got = h.prog_run(f"libprog_rast_dropin_{size}.so", script)
        want = h.prog_run(f"libprog_rast_ref_{size}.so", script)
        assert np.array_equal(got, want), (script, np.count_nonzero(got != want))


@pytest.mark.parametrize("setting,setting_boxes,W,H,f,cam_pos,yaw", [
    (2, 3, 320, 240, 120.0, h.DEFAULT_RAST_CAM, 0.0), (3, 2, 320, 240, 120.0, (0.1, -0.05, -2.6, 1.0), 0.174533),
    (2, 1, 900, 720, 512.0, h.DEFAULT_RAST_CAM, 0.0), (3, 3, 640, 480, 300.0, (0.0, 0.1, -1.6, 1.0), -0.349066)])
def test_textures_of_the_reference_repository(b200, renderer, setting, setting_boxes, W, H, f, cam_pos, yaw):
    """The metal-grill and woven-wood images of the reference's own repository (decoded pixels committed as
    tests/golden/rast_reference_textures.npz; the oracle is pinned on the compiled reference with the same images by
    tests/test_oracle_rast.py): real hole patterns and normal maps, incl. BASELINE config 2's resolution with the
    settings the reference ships with (2, 1)."""
    tex = h.reference_textures()
    room, boxes = cornell(setting, setting_boxes)
    renderer.set_textures(tex)
    try:
        check_draw(b200, renderer, tex, W, H, f, cam_pos, yaw, h.DEFAULT_RAST_LIGHT, room, boxes,
                   f"reference images, textures {setting}/{setting_boxes}", buffers=W <= 320)
    finally:
        renderer.set_textures(None)
        h.oracle_rast_set_textures(None)
