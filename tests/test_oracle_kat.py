"""Pins the CPU oracle (oracle/oracle.cpp).

The reference has no render tests, golden vectors or fixtures (SURVEY §4,
§8(c)): its whole suite is four Vec3 arithmetic tests, replayed first below.
Everything else here is a known-answer vector derived by hand (or by an
independent numpy restatement inside this file) from the reference lines cited
in each test, plus the published Philox4x32-10 known-answer vectors.  All
paths are relative to /root/reference/racer-tracer/.
"""
import ctypes as C
import math

import numpy as np
import pytest

from conftest import SCENES, scene_path
from racer_tracer_b200 import capi, harness


def d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


# ---------------------------------------------------------------------------
# The reference's own tests: src/vec3.rs:446-503
# ---------------------------------------------------------------------------
def vec_op(L, op, a, b):
    out = d3([0, 0, 0])
    L.oracle_vec3_op(op, d3(a), d3(b), out)
    return list(out)


def test_reference_vec3_add(oracle):       # vec3.rs:449-460
    L = oracle.lib()
    assert vec_op(L, 0, [1, 2, 3], [2, 4, 6]) == [3.0, 6.0, 9.0]
    assert vec_op(L, 0, [2, 4, 6], [1, 2, 3]) == [3.0, 6.0, 9.0]


def test_reference_vec3_sub(oracle):       # vec3.rs:462-475
    L = oracle.lib()
    assert vec_op(L, 1, [1, 2, 3], [2, 4, 6]) == [-1.0, -2.0, -3.0]
    assert vec_op(L, 1, [2, 4, 6], [1, 2, 3]) == [1.0, 2.0, 3.0]


def test_reference_vec3_mul(oracle):       # vec3.rs:477-492
    L = oracle.lib()
    assert vec_op(L, 2, [1, -2, 3], [5, 5, 5]) == [5.0, -10.0, 15.0]
    assert vec_op(L, 2, [1, -2, 3], [4, 8, 16]) == [4.0, -16.0, 48.0]


def test_reference_vec3_div(oracle):       # vec3.rs:494-502
    L = oracle.lib()
    assert vec_op(L, 3, [1, -2, 3], [2, 0, 0]) == [0.5, -1.0, 1.5]


# ---------------------------------------------------------------------------
# Philox4x32-10 published known-answer vectors (Random123 kat_vectors)
# ---------------------------------------------------------------------------
PHILOX_KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_known_answers(oracle, ctr, key, want):
    out = (C.c_uint32 * 4)()
    oracle.lib().oracle_philox4x32((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), 10, out)
    assert tuple(out) == want
    assert harness.philox4x32(ctr, key, 10) == want   # the host's copy (Perlin tables) agrees


PHILOX2_KAT = [   # Random123 kat_vectors, philox2x32 10
    ((0, 0), 0, (0xff1dae59, 0x6cd10df2)),
    ((0xffffffff, 0xffffffff), 0xffffffff, (0x2c3f628b, 0xab4fd7ad)),
    ((0x243f6a88, 0x85a308d3), 0x13198a2e, (0xdd7ce038, 0xf62a4c12)),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX2_KAT)
def test_philox2x32_known_answers(oracle, ctr, key, want):
    out = (C.c_uint32 * 2)()
    oracle.lib().oracle_philox2x32((C.c_uint32 * 2)(*ctr), key, 10, out)
    assert tuple(out) == want


# ---------------------------------------------------------------------------
# Geometry
# ---------------------------------------------------------------------------
def mini_scene(prims):
    """prims: list of (type, data5) -> FlatScene with one lambertian material."""
    fs = harness.FlatScene()
    n = len(prims)
    fs.a_type = np.array([p[0] for p in prims], dtype=np.int32)
    fs.a_data = np.array([p[1] for p in prims], dtype=np.float64).reshape(n, 5)
    fs.a_mat = np.zeros(n, dtype=np.int32)
    fs.a_id = np.arange(1, n + 1, dtype=np.uint32) << 3
    fs.a_aabb = np.zeros((n, 6))
    c = fs.c
    c.n_prims = n
    c.prim_type = fs.a_type.ctypes.data_as(C.POINTER(C.c_int32))
    c.prim_data = fs.a_data.ctypes.data_as(C.POINTER(C.c_double))
    c.prim_material = fs.a_mat.ctypes.data_as(C.POINTER(C.c_int32))
    c.prim_id = fs.a_id.ctypes.data_as(C.POINTER(C.c_uint32))
    c.prim_aabb = fs.a_aabb.ctypes.data_as(C.POINTER(C.c_double))
    fs.mats = (capi.rc_material * 1)()
    fs.texs = (capi.rc_texture * 1)()
    fs.texs[0].color[:] = [0.5, 0.5, 0.5]
    c.n_materials, c.materials = 1, fs.mats
    c.n_textures, c.textures = 1, fs.texs
    c.bg_type = capi.RC_BG_SKY
    c.bg_a[:] = [1.0, 1.0, 1.0]
    c.bg_b[:] = [0.5, 0.7, 1.0]
    return fs


def prim_hit(L, fs, i, o, d, t_min=0.001, t_max=math.inf):
    out = (C.c_double * 10)()
    ok = L.oracle_prim_hit(fs.ptr, i, d3(o), d3(d), t_min, t_max, out)
    return (ok, list(out))


def test_sphere_hit_known_answers(oracle):
    """src/geometry/sphere.rs:31-68 + get_sphere_uv :20-27."""
    L = oracle.lib()
    fs = mini_scene([(capi.RC_PRIM_SPHERE, [0, 0, -1, 0.5, 0])])
    ok, r = prim_hit(L, fs, 0, [0, 0, 0], [0, 0, -1])
    assert ok and r[0] == 0.5 and r[1:4] == [0, 0, -0.5] and r[4:7] == [0, 0, 1] and r[9] == 1.0
    # uv of outward normal (0,0,1): theta = acos(0) = pi/2, phi = atan2(-1, 0) + pi = pi/2
    assert r[7] == pytest.approx(0.25, abs=1e-15) and r[8] == pytest.approx(0.5, abs=1e-15)
    # unnormalised direction (Q3): t scales with 1/|d|
    ok, r = prim_hit(L, fs, 0, [0, 0, 0], [0, 0, -4])
    assert ok and r[0] == 0.125
    # from inside: near root negative -> far root, normal flipped against the ray, front_face false
    ok, r = prim_hit(L, fs, 0, [0, 0, -1], [1, 0, 0])
    assert ok and r[0] == 0.5 and r[4:7] == [-1, 0, 0] and r[9] == 0.0
    # t_max is inclusive (sphere.rs:53: `t_max < root` rejects)
    assert prim_hit(L, fs, 0, [0, 0, 0], [0, 0, -1], t_max=0.5)[0] == 1
    assert prim_hit(L, fs, 0, [0, 0, 0], [0, 0, -1], t_max=0.4999)[0] == 0
    # miss
    assert prim_hit(L, fs, 0, [0, 1, 0], [0, 0, -1])[0] == 0
    # negative radius flips the outward normal (Q20, three_balls.yml dialectric_inner)
    fs2 = mini_scene([(capi.RC_PRIM_SPHERE, [0, 0, -1, -0.5, 0])])
    ok, r = prim_hit(L, fs2, 0, [0, 0, 0], [0, 0, -1])
    assert ok and r[0] == 0.5 and r[4:7] == [0, 0, 1] and r[9] == 0.0   # outward = -z, not against the ray


def test_rect_hit_known_answers(oracle):
    """src/geometry/xy_rect.rs:21-48, xz_rect.rs:21-49, yz_rect.rs:21-49."""
    L = oracle.lib()
    fs = mini_scene([(capi.RC_PRIM_XY_RECT, [0, 2, 0, 4, -3]),
                     (capi.RC_PRIM_XZ_RECT, [0, 2, 0, 4, 5]),
                     (capi.RC_PRIM_YZ_RECT, [0, 2, 0, 4, 7])])
    ok, r = prim_hit(L, fs, 0, [0.5, 1, 0], [0, 0, -1.5])
    assert ok and r[0] == 2.0 and r[1:4] == [0.5, 1, -3] and r[4:7] == [0, 0, 1] and r[7:9] == [0.25, 0.25]
    ok, r = prim_hit(L, fs, 0, [0.5, 1, -6], [0, 0, 1])      # from behind: normal flips
    assert ok and r[4:7] == [0, 0, -1] and r[9] == 0.0
    ok, r = prim_hit(L, fs, 1, [1, 0, 1], [0, 10, 0])
    assert ok and r[0] == 0.5 and r[4:7] == [0, -1, 0] and r[7:9] == [0.5, 0.25]
    ok, r = prim_hit(L, fs, 2, [0, 1, 3], [14, 0, 0])
    assert ok and r[0] == 0.5 and r[4:7] == [-1, 0, 0] and r[7:9] == [0.5, 0.75]
    # bounds are inclusive (`x < x0 || x > x1` rejects)
    assert prim_hit(L, fs, 0, [2, 4, 0], [0, 0, -1])[0] == 1
    assert prim_hit(L, fs, 0, [2.0000001, 4, 0], [0, 0, -1])[0] == 0
    # t range
    assert prim_hit(L, fs, 0, [0.5, 1, 0], [0, 0, -1], t_max=2.999)[0] == 0
    assert prim_hit(L, fs, 0, [0.5, 1, -3.0005], [0, 0, 1])[0] == 0   # t = 0.0005 < t_min


def test_aabb_hit_is_the_reference_per_axis_test(oracle):
    """src/aabb.rs:42-59 clips every axis against the ORIGINAL interval (Q12):
    a ray whose x-slab is t in [1,2] and y-slab t in [3,4] still 'hits'."""
    L = oracle.lib()
    lo, hi = d3([0, 0, 0]), d3([1, 1, 1])
    assert L.oracle_aabb_hit(lo, hi, d3([-1, 4, 0.5]), d3([1, -1, 0]), 0.001, math.inf) == 1
    assert L.oracle_aabb_hit(lo, hi, d3([-1, 0.5, 0.5]), d3([1, 0, 0]), 0.001, math.inf) == 1
    assert L.oracle_aabb_hit(lo, hi, d3([-1, 0.5, 0.5]), d3([-1, 0, 0]), 0.001, math.inf) == 0   # behind
    assert L.oracle_aabb_hit(lo, hi, d3([-1, 0.5, 0.5]), d3([1, 0, 0]), 0.001, 1.0) == 0         # max <= min
    assert L.oracle_aabb_hit(lo, hi, d3([-1, 2.5, 0.5]), d3([1, 0, 0]), 0.001, math.inf) == 0    # y outside, dy = 0


def test_linear_list_tie_goes_to_the_later_primitive(oracle):
    """src/shared_scene.rs:37-53 with `t <= closest` accepted (sphere.rs:53)."""
    L = oracle.lib()
    fs = mini_scene([(capi.RC_PRIM_XY_RECT, [0, 2, 0, 2, -1]), (capi.RC_PRIM_XY_RECT, [0, 2, 0, 2, -1])])
    out = (C.c_double * 10)()
    assert L.oracle_scene_hit(fs.ptr, d3([1, 1, 0]), d3([0, 0, -1]), 0.001, math.inf, out) == 1


# ---------------------------------------------------------------------------
# Materials / vector helpers
# ---------------------------------------------------------------------------
def test_reflectance_reflect_refract(oracle):
    L = oracle.lib()
    # dialectric.rs:17-22 (Schlick): r0 = ((1-1.5)/(1+1.5))^2 = 0.04
    assert L.oracle_reflectance(1.0, 1.5) == pytest.approx(0.04, abs=1e-16)
    assert L.oracle_reflectance(0.0, 1.5) == pytest.approx(1.0, abs=1e-16)
    assert L.oracle_reflectance(0.5, 1.5) == pytest.approx(0.04 + 0.96 * 0.5 ** 5, abs=1e-16)
    out = d3([0, 0, 0])
    L.oracle_reflect(d3([1, -1, 0]), d3([0, 1, 0]), out)            # vec3.rs:412-414
    assert list(out) == [1.0, 1.0, 0.0]
    # vec3.rs:416-422: 45 degrees into glass, eta = 1/1.5
    s = math.sqrt(0.5)
    L.oracle_refract(d3([s, -s, 0]), d3([0, 1, 0]), 1 / 1.5, out)
    perp_x = (1 / 1.5) * s
    assert out[0] == pytest.approx(perp_x, abs=1e-15)
    assert out[1] == pytest.approx(-math.sqrt(1 - perp_x ** 2), abs=1e-15)
    assert out[0] ** 2 + out[1] ** 2 == pytest.approx(1.0, abs=1e-15)   # Snell: sin(t) = sin(i)/1.5


def test_background_sky_and_solid(oracle):
    """src/background_color.rs:27-48: (1-t)*top + t*bottom, t = 0.5*(unit.y + 1)."""
    L = oracle.lib()
    fs = mini_scene([])
    out = d3([0, 0, 0])
    L.oracle_background(fs.ptr, d3([0, 5, 0]), out)
    assert list(out) == [0.5, 0.7, 1.0]
    L.oracle_background(fs.ptr, d3([0, -2, 0]), out)
    assert list(out) == [1.0, 1.0, 1.0]
    L.oracle_background(fs.ptr, d3([3, 0, 0]), out)
    assert list(out) == pytest.approx([0.75, 0.85, 1.0], abs=1e-15)
    fs.c.bg_type = capi.RC_BG_SOLID
    fs.c.bg_a[:] = [0.1, 0.2, 0.3]
    L.oracle_background(fs.ptr, d3([3, 1, 0]), out)
    assert list(out) == [0.1, 0.2, 0.3]


# ---------------------------------------------------------------------------
# Textures
# ---------------------------------------------------------------------------
def np_perlin_noise(p, pt):
    """independent numpy restatement of src/texture/noise.rs:57-96"""
    g = np.array([[p.ran_vec[i][k] for k in range(3)] for i in range(256)])
    f = np.floor(pt)
    u, v, w = pt - f
    i, j, k = (int(x) for x in f)
    uu, vv, ww = (x * x * (3 - 2 * x) for x in (u, v, w))
    acc = 0.0
    for a in range(2):
        for b in range(2):
            for c in range(2):
                idx = p.perm_x[(i + a) & 255] ^ p.perm_y[(j + b) & 255] ^ p.perm_z[(k + c) & 255]
                acc += ((a * uu + (1 - a) * (1 - uu)) * (b * vv + (1 - b) * (1 - vv)) *
                        (c * ww + (1 - c) * (1 - ww)) * float(g[idx] @ np.array([u - a, v - b, w - c])))
    return acc


def test_perlin_noise_and_turbulence(oracle):
    L = oracle.lib()
    p = harness.make_perlin(seed=0, index=0)
    # identity permutations (noise.rs:122 iterates an empty range, Q17) and unit gradients
    assert list(p.perm_x) == list(range(256)) == list(p.perm_y) == list(p.perm_z)
    g = np.array([[p.ran_vec[i][k] for k in range(3)] for i in range(256)])
    assert np.allclose(np.linalg.norm(g, axis=1), 1.0, atol=1e-15)
    assert L.oracle_perlin_noise(C.byref(p), d3([3, -7, 11])) == 0.0           # lattice points are zeros
    rng = np.random.default_rng(1)
    for pt in rng.uniform(-40, 40, size=(20, 3)):
        assert L.oracle_perlin_noise(C.byref(p), d3(pt)) == pytest.approx(np_perlin_noise(p, pt), abs=1e-13)
    pt = np.array([1.3, -2.6, 0.77])
    turb = abs(sum(0.5 ** o * np_perlin_noise(p, pt * 2 ** o) for o in range(7)))   # noise.rs:98-109
    assert L.oracle_perlin_turbulence(C.byref(p), d3(pt), 7) == pytest.approx(turb, abs=1e-13)


def test_texture_values(oracle, cfg):
    L = oracle.lib()
    job = harness.prepare_job(scene_path("noise_and_textures"), cfg, 64, 64)
    fs = job.scene
    kinds = {fs.textures[i].type: i for i in range(len(fs.textures))}
    out = d3([0, 0, 0])
    # checker: sin(10x) sin(10y) sin(10z) < 0 -> odd (texture_b) (checkered.rs:32-43)
    ck = kinds[capi.RC_TEX_CHECKER]
    for pt in ([0.1, 0.1, 0.1], [0.1, 0.1, -0.1], [2.3, -0.4, 5.5]):
        L.oracle_texture_value(fs.ptr, ck, 0.0, 0.0, d3(pt), out)
        odd = math.sin(10 * pt[0]) * math.sin(10 * pt[1]) * math.sin(10 * pt[2]) < 0
        want = fs.textures[fs.textures[ck].b].color if odd else fs.textures[fs.textures[ck].a].color
        assert list(out) == list(want)
    # image: nearest texel, v flipped, index clamped (image.rs:28-51)
    im = kinds[capi.RC_TEX_IMAGE]
    w, h, px = fs.images[0]
    for (u, v) in [(0.0, 1.0), (1.0, 0.0), (0.5, 0.5), (0.2501, 0.7499), (-3.0, 7.0)]:
        L.oracle_texture_value(fs.ptr, im, u, v, d3([0, 0, 0]), out)
        uc, vc = min(max(u, 0.0), 1.0), 1.0 - min(max(v, 0.0), 1.0)
        i, j = min(int(uc * w), w - 1), min(int(vc * h), h - 1)
        assert list(out) == [px[j, i, c] * (1.0 / 255.0) for c in range(3)]
    # noise: color * 0.5 * (1 + sin(scale*z + 10*turb)) (noise.rs:26-33)
    nz = kinds[capi.RC_TEX_NOISE]
    pt = [0.4, 1.9, -0.3]
    L.oracle_texture_value(fs.ptr, nz, 0.0, 0.0, d3(pt), out)
    turb = L.oracle_perlin_turbulence(C.byref(fs.c.perlin[0]), d3(pt), 7)
    assert out[0] == pytest.approx(0.5 * (1 + math.sin(4 * pt[2] + 10 * turb)), abs=1e-15)


# ---------------------------------------------------------------------------
# Camera
# ---------------------------------------------------------------------------
def test_camera_derivation_and_get_ray(oracle, cfg):
    """src/camera.rs:196-234 and :326-337; cornell_box.yml camera."""
    L = oracle.lib()
    cam = capi.rc_camera()
    L.oracle_camera_new(d3([278, 278, -800]), d3([278, 278, 0]), d3([0, 1, 0]), 40.0, 0.0, 10000.0, 16 / 9,
                        0.0, 1.0, C.byref(cam))
    h = math.tan(math.radians(40) / 2)
    assert list(cam.forward) == [0, 0, -1] and list(cam.right) == [-1, 0, 0] and list(cam.up) == [0, 1, 0]
    assert cam.viewport_height == 2 * h and cam.viewport_width == 16 / 9 * 2 * h
    assert list(cam.horizontal) == [-10000 * cam.viewport_width, 0, 0]
    assert list(cam.vertical) == [0, 10000 * cam.viewport_height, 0]
    assert cam.upper_left_corner[2] == pytest.approx(-800 + 10000)
    mine = harness.make_camera(harness.merged_camera(None, cfg.camera), 1920, 1080)
    for name, _ in capi.rc_camera._fields_:
        a, b = getattr(cam, name), getattr(mine, name)
        assert (list(a) == list(b)) if hasattr(a, "__len__") else (a == b), name
    o, d = d3([0, 0, 0]), d3([0, 0, 0])
    L.oracle_get_ray(C.byref(cam), 0.5, 0.5, 0.0, 0.0, o, d)        # centre of the frame looks down +z
    assert list(o) == [278, 278, -800] and d[0] == pytest.approx(0, abs=1e-9) and d[2] == 10000.0
    # lens offset moves the origin and is subtracted from the direction (camera.rs:328-334)
    cam.lens_radius = 2.0
    L.oracle_get_ray(C.byref(cam), 0.5, 0.5, 0.5, -0.25, o, d)
    assert list(o) == [278 - 1.0, 278 - 0.5, -800]
    assert d[0] == pytest.approx(1.0, abs=1e-9) and d[1] == pytest.approx(0.5, abs=1e-9)


# ---------------------------------------------------------------------------
# Tone maps + quantiser
# ---------------------------------------------------------------------------
def test_tone_maps_against_numpy(oracle):
    rng = np.random.default_rng(0)
    img = rng.random((5, 7, 3)) * 2.0
    # ACES: out_matrix * fit(in_matrix * c) (aces.rs:18-55, tone_map.rs:45-60)
    tm = harness.make_tone_map({"aces": {"default": True}})
    mi = np.array(list(tm.aces_in)).reshape(3, 3)
    mo = np.array(list(tm.aces_out)).reshape(3, 3)
    x = img @ mi.T
    fit = (x * (x + 0.0245786) - 0.000090537) / (x * (0.983729 * x + 0.4329510) + 0.238081)
    assert np.allclose(oracle.tone_map(tm, img), fit @ mo.T, rtol=1e-13)
    # Reinhard extended luminance, max_white 25 -> squared (reinhard.rs:10-42)
    tm = harness.make_tone_map({"reinhard": {"default": True}})
    l_old = img @ np.array([0.2126, 0.7152, 0.0722])
    l_new = l_old * (1 + l_old / 625.0) / (1 + l_old)
    assert np.allclose(oracle.tone_map(tm, img), img * (l_new / l_old)[..., None], rtol=1e-13)
    assert np.isnan(oracle.tone_map(tm, np.zeros((1, 1, 3)))).all()     # 0/0, as the reference
    # Hable filmic (hable.rs:41-80), defaults tone_map.rs:24-44
    tm = harness.make_tone_map({"hable": {"default": True}})
    A, B, Cc, D, E, F = 0.15, 0.5, 0.1, 0.2, 0.02, 0.3
    part = lambda v: ((v * (A * v + Cc * B) + D * E) / (v * (A * v + B) + D * F)) - E / F
    assert np.allclose(oracle.tone_map(tm, img), part(img * 2.0) * (1.0 / part(11.2)), rtol=1e-13)
    tm = harness.make_tone_map("None")
    assert np.array_equal(oracle.tone_map(tm, img), img)


def test_quantiser_has_no_clamp(oracle):
    """src/image_action/png.rs:21-31: (v*255) as u32, (r<<24)|(g<<16)|(b<<8)|255, big-endian bytes."""
    q = oracle.quantise_rgba(np.array([[[0.0, 0.5, 1.0], [1.5, 0.5, 0.2], [-1.0, float("nan"), 0.999]]]))
    assert tuple(q[0, 0]) == (0, 127, 255, 255)
    # r = 382 = 0x17E: bit 8 falls off the top of the u32; g = 127, b = 51
    assert tuple(q[0, 1]) == (0x7E, 0x7F, 0x33, 0xFF)
    # negative and NaN saturate to 0 (Rust `as u32`)
    assert tuple(q[0, 2]) == (0, 0, 254, 255)
    # green > 255 bleeds into red
    q = oracle.quantise_rgba(np.array([[[0.0, 2.0, 0.0]]]))
    assert tuple(q[0, 0]) == (0x01, 0xFE, 0x00, 0xFF)


# ---------------------------------------------------------------------------
# ray_color semantics (src/renderer.rs:41-90) and the pixel loop (cpu.rs:26-71)
# ---------------------------------------------------------------------------
def test_depth_exhaustion_is_white(oracle, cfg):
    job = harness.prepare_job(scene_path("cornell_box"), cfg, 16, 16)
    img = oracle.render(job, harness.make_params(16, 16, 3, 0))
    assert np.array_equal(img, np.ones_like(img))                      # renderer.rs:48-56


def test_single_bounce_is_emission_or_background_plus_white(oracle, cfg):
    """max_depth = 1: a light returns its emission, a miss the background, and any
    scattering surface emitted(0) + attenuation * white."""
    job = harness.prepare_job(scene_path("cornell_box"), cfg, 32, 32)
    p = harness.make_params(32, 32, 1, 1, fixed_jitter=1)
    lin = oracle.render(job, p, linear_sum=True)
    ids, _, _, _ = oracle.primary_aov(job, p)
    ids = ids.reshape(32, 32)
    keys = job.scene.object_keys
    colour = {"piece_1": [0.12, 0.45, 0.15], "piece_2": [0.65, 0.05, 0.05], "piece_3": [0.63] * 3,
              "piece_4": [0.63] * 3, "piece_5": [0.63] * 3, "light": [15.0] * 3}
    assert (ids == 0).any()
    for y in range(32):
        for x in range(32):
            want = [0.0, 0.0, 0.0] if ids[y, x] == 0 else colour[keys[(ids[y, x] >> 3) - 1]]
            assert list(lin[y, x]) == want


def test_tile_grid_does_not_change_the_image(oracle, cfg):
    """prepare_threads (cpu.rs:73-115): remainder columns/rows go to the last tile;
    with counter-based streams the image is independent of the tile grid."""
    job = harness.prepare_job(scene_path("three_balls"), cfg, 47, 31)
    p = harness.make_params(47, 31, 2, 5, seed=4)
    a = oracle.render(job, p, tiles=(10, 10))
    b = oracle.render(job, p, tiles=(1, 1), threads=1)
    c = oracle.render(job, p, tiles=(7, 3), threads=3)
    assert np.array_equal(a, b) and np.array_equal(a, c)


def test_u_jitter_is_per_pixel_v_jitter_per_sample(oracle, cfg):
    """cpu.rs:35-36 vs :39-40 (Q1): with a vertical edge in view the per-pixel u
    jitter makes every sample of a pixel see the same column."""
    job = harness.prepare_job(scene_path("cornell_box"), cfg, 64, 36)
    p = harness.make_params(64, 36, 64, 1, seed=5)
    lin = oracle.render(job, p, linear_sum=True) / 64.0
    # depth 1 radiance is a pure function of the primary hit: walls give their albedo,
    # so horizontally every pixel is ONE colour (no blend between the black background and the wall)
    row = lin[18]
    allowed = [[0, 0, 0], [0.12, 0.45, 0.15], [0.65, 0.05, 0.05], [0.63, 0.63, 0.63]]
    for px in row:
        assert any(np.allclose(px, a, atol=1e-12) for a in allowed), px


def test_sample_ranges_add_up(oracle, cfg):
    job = harness.prepare_job(scene_path("emissive"), cfg, 24, 24)
    p = harness.make_params(24, 24, 8, 20, seed=6)
    whole = oracle.render(job, p, linear_sum=True)
    a = oracle.render(job, p, linear_sum=True, sample_begin=0, sample_count=5)
    b = oracle.render(job, p, linear_sum=True, sample_begin=5, sample_count=3)
    assert np.allclose(a + b, whole, rtol=1e-13, atol=1e-13)


def test_samplers_and_rng_back_ends_agree_statistically(oracle, cfg):
    """The direct (inverse-transform) samplers draw the same distributions as the
    reference's rejection loops (vec3.rs:424-444, util.rs:25-39); the Philox and
    the sequential back ends are both uniform.  Mean radiance must agree within
    Monte-Carlo error."""
    job = harness.prepare_job(scene_path("three_balls"), cfg, 40, 40)
    means = []
    for sampler, rng in [(capi.RC_SAMPLER_REJECTION, oracle.RNG_SEQUENTIAL), (capi.RC_SAMPLER_DIRECT, oracle.RNG_SEQUENTIAL),
                         (capi.RC_SAMPLER_REJECTION, oracle.RNG_PHILOX), (capi.RC_SAMPLER_DIRECT, oracle.RNG_PHILOX)]:
        p = harness.make_params(40, 40, 256, 20, seed=21, sampler=sampler)
        lin = oracle.render(job, p, rng=rng, linear_sum=True) / 256.0
        means.append(lin.mean(axis=(0, 1)))
    means = np.array(means)
    assert np.abs(means - means[0]).max() < 0.004, means


def test_counters_are_consistent(oracle, cfg):
    job = harness.prepare_job(scene_path("cornell_box"), cfg, 64, 36)
    p = harness.make_params(64, 36, 16, 20, seed=1)
    _, cnt = oracle.render(job, p, want_counters=True)
    d = cnt.as_dict()
    assert d["samples"] == 64 * 36 * 16
    # every segment ends in a background miss or a primitive hit
    assert d["segments"] == d["background"] + sum(d["prim_hits"]) - (sum(d["prim_hits"]) - sum(d["scatters"]))
    assert sum(d["scatters"]) + d["background"] == d["segments"]
    assert d["node_tests"] >= d["segments"]
    assert 2.5 < d["segments"] / d["samples"] < 4.5          # ~3.5 segments per sample


@pytest.mark.parametrize("name", SCENES)
def test_bvh_and_linear_list_give_the_same_image(oracle, cfg, name):
    """Node::hit (bvh_node.rs:112-132) vs the linear list (shared_scene.rs:37-53)."""
    # 53x31: not square, so no pixel-centre ray runs exactly along the Cornell box's 45-degree
    # wall/floor edges, where the two traversal orders break the exact tie differently (Q13)
    a = harness.prepare_job(scene_path(name), cfg, 53, 31, use_bvh=True)
    b = harness.prepare_job(scene_path(name), cfg, 53, 31, use_bvh=False)
    p = harness.make_params(53, 31, 2, 20, seed=8)
    ia, ib = oracle.render(a, p), oracle.render(b, p)
    assert np.array_equal(ia, ib)
    assert np.array_equal(oracle.primary_aov(a, p)[0], oracle.primary_aov(b, p)[0])


def test_preview_renderer_semantics(cfg):
    """CpuRendererScaled (src/renderer/cpu_scaled.rs): scale derivation (:17-43), one colour per block
    (:75-86), remainder pixels of a tile left at Vec3::default() (:53), scale 1 == the full renderer."""
    from oracle import oracle as O
    assert harness.preview_scales(cfg, 600, 600) == (4, 4)
    assert harness.preview_scales(cfg, 610, 330) == (1, 3)
    assert harness.preview_scales(cfg, 1920, 1080) == (4, 4)
    w, h = 50, 35
    job = harness.prepare_job(scene_path("three_balls"), cfg, w, h)
    p = harness.make_params(w, h, 4, 5, seed=1)
    # 5x5 tiles of 10x7 pixels, 3x3 blocks: 3 whole blocks per tile row (9 of 10 columns), 2 per column (6 of 7 rows)
    img = O.render_preview(job, p, 3, 3, tiles=(5, 5))
    tile = img[:7, :10]
    assert tile[:6, :9].any() and not tile[6:].any() and not tile[:, 9:].any()
    assert np.array_equal(tile[0:3, 0:3], np.broadcast_to(tile[0, 0], (3, 3, 3)))
    one = O.render_preview(job, p, 1, 1, tiles=(5, 5))
    assert np.allclose(one, O.render(job, p, tiles=(5, 5)), rtol=0, atol=1e-12)


def test_moving_sphere_and_random_scene(cfg):
    """MovingSphere (src/geometry/moving_sphere.rs): centre = pos + (time - ta)/(tb - ta) (pos_b - pos), the
    Sphere quadratic around it, box = union of both ends; Random::load (src/scene/random.rs) builds ~480
    spheres of which the diffuse ones move; its camera has an aperture, so the lens is sampled."""
    from oracle import oracle as O
    job = harness.prepare_job("random", cfg, 64, 48, seed=3)
    fs = job.scene
    types = fs.np["prim_type"]
    assert fs.c.n_prims > 400 and fs.c.n_nodes == 2 * fs.c.n_prims - 1
    mov = np.flatnonzero(types == capi.RC_PRIM_MOVING_SPHERE)
    assert 0.7 < len(mov) / (fs.c.n_prims - 4) < 0.9                   # 80 % of the small spheres
    i = int(mov[0])
    pos, r = fs.np["prim_data"][i, :3], fs.np["prim_data"][i, 3]
    pos_b, ta, tb = fs.np["prim_motion"][i, :3], fs.np["prim_motion"][i, 3], fs.np["prim_motion"][i, 4]
    assert (ta, tb) == (0.0, 1.0) and pos_b[0] == pos[0] and pos_b[2] == pos[2] and 0.0 <= pos_b[1] - pos[1] < 0.5
    box = fs.np["prim_aabb"][i]
    assert np.allclose(box[:3], np.minimum(pos, pos_b) - r) and np.allclose(box[3:], np.maximum(pos, pos_b) + r)
    # a ray straight down onto the sphere hits its top at the centre of that ray time; fixed jitter = time_a
    out = (C.c_double * 10)()
    org = O.d3([pos[0], pos[1] + 5.0, pos[2]])
    assert O.lib().oracle_prim_hit(fs.ptr, i, org, O.d3([0.0, -1.0, 0.0]), 0.001, 1e30, out) == 1
    assert abs(out[0] - (5.0 - r)) < 1e-12 and abs(out[5] - 1.0) < 1e-12   # t, normal.y at time 0
    # the generator is deterministic in the seed and differs between seeds
    again = harness.prepare_job("random", cfg, 64, 48, seed=3).scene
    other = harness.prepare_job("random", cfg, 64, 48, seed=4).scene
    assert np.array_equal(again.np["prim_data"], fs.np["prim_data"]) and not np.array_equal(other.np["prim_data"][:8], fs.np["prim_data"][:8])
    assert job.camera.lens_radius == 0.05 and job.camera.focus_distance == 10.0
    # motion blur: a moving sphere's pixels differ between time-sampled and time-frozen renders
    p = harness.make_params(64, 48, 4, 5, seed=1)
    img = O.render(job, p)
    assert np.isfinite(img).all() and img.std() > 0.05


def test_the_tree_does_not_change_the_image(oracle, cfg):
    """Random scene under the three structures a host may hand over — the SAH tree the hosts build for large
    scenes, the reference's median split (bvh_node.rs:31-82) and no tree at all (the linear list,
    shared_scene.rs:37-53): the closest hit of every ray, hence every primary id and every pixel, is the same."""
    sah = harness.prepare_job("random", cfg, 53, 31, seed=0)
    saved = harness.SAH_MIN_OBJECTS
    try:
        harness.SAH_MIN_OBJECTS = 1 << 30
        median = harness.prepare_job("random", cfg, 53, 31, seed=0)
    finally:
        harness.SAH_MIN_OBJECTS = saved
    linear = harness.prepare_job("random", cfg, 53, 31, seed=0, use_bvh=False)
    assert [n.left for n in sah.scene.nodes] != [n.left for n in median.scene.nodes] and linear.scene.c.n_nodes == 0
    p = harness.make_params(53, 31, 2, 20, seed=8)
    ids = oracle.primary_aov(sah, p)[0]
    img = oracle.render(sah, p)
    for other in (median, linear):
        assert np.array_equal(oracle.primary_aov(other, p)[0], ids)
        assert np.array_equal(oracle.render(other, p), img)
