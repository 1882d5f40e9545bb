"""The C++ host above the C ABI (racer_tracer_b200/host/): the mirror of the reference's Rust host for
this path — YAML scene/config loading (src/scene/yml.rs, src/config.rs), camera derivation
(src/camera.rs), flattening + host BVH (src/bvh_node.rs), the Renderer trait (src/renderer.rs:92-116),
ScreenBuffer tone map and SavePng (src/image_buffer.rs, src/image_action/png.rs).

CPU tests compare what the C++ host hands to the C ABI with what the Python harness (which feeds the
oracle) hands to it, field for field; the GPU test renders through `racer_render` and compares the PNG
bytes with the harness path."""
import ctypes as C
import json
import os
import struct
import subprocess
import zlib

import numpy as np
import pytest

from conftest import ROOT, SCENES_X, scene_path
from racer_tracer_b200 import capi, harness

PKG = os.path.join(ROOT, "racer_tracer_b200")
BIN = os.path.join(PKG, "racer_render")
CONFIG = os.path.join(ROOT, "tests", "golden", "config.yml")
IMAGES = os.path.join(ROOT, "tests", "golden", "resources", "images")


@pytest.fixture(scope="session")
def racer_render():
    if not os.path.exists(BIN) or os.path.getmtime(BIN) < os.path.getmtime(os.path.join(PKG, "host", "racer_host.cpp")):
        subprocess.run(["bash", os.path.join(PKG, "host", "build.sh")], check=True)
    return BIN


def run(racer_render, *args, check=True):
    r = subprocess.run([racer_render, *args], capture_output=True, text=True)
    if check:
        assert r.returncode == 0, r.stderr
    return r


def dump(racer_render, tmp_path, scene, *extra):
    out = tmp_path / "flat.json"
    run(racer_render, "--config", CONFIG, "--scene", scene, "--image-dir", IMAGES, "--dump-flat", str(out), *extra)
    return json.loads(out.read_text())


@pytest.mark.parametrize("name", SCENES_X + ["two_balls"])
def test_flattened_scene_equals_the_python_harness(racer_render, tmp_path, cfg, name):
    d = dump(racer_render, tmp_path, scene_path(name))
    job = harness.prepare_job(scene_path(name), cfg, image_dirs=[IMAGES])
    fs, s = job.scene, d["scene"]
    assert s["prim_type"] == fs.np["prim_type"].tolist()
    assert s["prim_material"] == fs.np["prim_material"].tolist()
    assert s["prim_id"] == fs.np["prim_id"].tolist()
    assert s["prim_instance"] == fs.np["prim_instance"].tolist()
    assert s["prim_data"] == fs.np["prim_data"].ravel().tolist()          # bit-exact f64
    assert s["prim_aabb"] == fs.np["prim_aabb"].ravel().tolist()
    assert s["prim_motion"] == fs.np["prim_motion"].ravel().tolist()
    assert s["object_keys"] == fs.object_keys
    assert [(m["type"], m["texture"], m["param"][0]) for m in s["materials"]] == [(m.type, m.texture, m.param) for m in fs.materials]
    assert [(t["type"], t["a"], t["b"], t["v"]) for t in s["textures"]] == \
        [(t.type, t.a, t.b, [t.color[0], t.color[1], t.color[2], t.scale]) for t in fs.textures]
    assert [(n["left"], n["right"], n["bmin"], n["bmax"]) for n in s["nodes"]] == \
        [(n.left, n.right, list(n.bmin), list(n.bmax)) for n in fs.nodes]
    assert [(i["flags"], i["v"]) for i in s["instances"]] == \
        [(i.flags, [i.sin_theta, i.cos_theta, i.offset[0], i.offset[1], i.offset[2]]) for i in fs.instances]
    assert (s["bg_type"], s["bg_a"], s["bg_b"]) == (fs.c.bg_type, list(fs.c.bg_a), list(fs.c.bg_b))
    if fs.c.n_perlin:
        assert s["perlin0"] == np.ctypeslib.as_array(fs.c.perlin[0].ran_vec).ravel().tolist()
    # camera (Camera::new) and tone map selection (main.rs:84-86), render / preview configuration
    cam = np.frombuffer(bytes(job.camera), dtype=np.float64).tolist()
    assert d["camera"] == cam
    assert d["tone_map_type"] == job.tone_map.type
    assert (d["width"], d["height"]) == (cfg.width, cfg.height)
    assert d["render"] == [cfg.render.samples, cfg.render.max_depth]
    assert d["preview"] == [cfg.preview.samples, cfg.preview.max_depth, cfg.preview.scale]


def test_baseline_jpeg_decoder_against_pillow(racer_render, tmp_path):
    """TextureImage::try_new decodes earthmap.jpg (texture/image.rs:17-25).  The C++ host has its own
    baseline decoder; decoders may differ by IDCT rounding, so: same size, every channel within 3 levels,
    mean absolute difference below 0.5 level."""
    d = dump(racer_render, tmp_path, scene_path("noise_and_textures"))
    from PIL import Image
    im = np.asarray(Image.open(os.path.join(IMAGES, "earthmap.jpg")).convert("RGBA"), dtype=np.int64)
    meta = d["scene"]["images"][0]
    assert (meta["width"], meta["height"]) == (im.shape[1], im.shape[0])
    assert abs(meta["byte_sum"] - int(im.sum())) / im.size < 0.5


def test_malformed_jpeg_is_an_error_not_a_crash(racer_render, tmp_path):
    """The reference decodes through the memory-safe `image` crate; the host's own decoder must reject what it
    cannot index: entropy-table selectors above 3 in the scan header, a scan component that is not in the frame, a
    truncated file.  Exit code 21 = TracerError::FailedToOpenImage (src/error.rs:71-97), never a signal."""
    good = open(os.path.join(IMAGES, "earthmap.jpg"), "rb").read()
    sos = good.index(b"\xff\xda")
    ns = good[sos + 4]
    assert ns == 3
    bad_selector = bytearray(good); bad_selector[sos + 6] = 0x4F            # component 1: td = 4, ta = 15
    bad_component = bytearray(good); bad_component[sos + 5] = 0x77          # an id the frame header does not list
    cases = {"selector.jpg": bytes(bad_selector), "component.jpg": bytes(bad_component), "truncated.jpg": good[:sos + 40]}
    scene = open(scene_path("noise_and_textures")).read()
    for name, data in cases.items():
        d = tmp_path / name.split(".")[0]
        (d / "scenes").mkdir(parents=True)
        (d / "resources" / "images").mkdir(parents=True)
        (d / "resources" / "images" / "earthmap.jpg").write_bytes(data)
        (d / "scenes" / "scene.yml").write_text(scene)
        r = run(racer_render, "--config", CONFIG, "--scene", str(d / "scenes" / "scene.yml"), "--dump-flat", str(d / "flat.json"), check=False)
        assert r.returncode == 21, (name, r.returncode, r.stderr[-300:])


def test_reference_style_yaml_and_errors(racer_render, tmp_path):
    """The reference's own files use `---`, values on the next line and spaced flow lists
    (resources/scenes/*.yml, config.yml:28-29); errors map to the reference's TracerError ordinals
    (src/error.rs:71-97)."""
    cfg_file = tmp_path / "config.yml"
    cfg_file.write_text("""preview:
  samples: 4
  max_depth: 3
  scale: 4
  num_threads_width: 10
  num_threads_height: 10

render:
  samples: 9
  max_depth: 7
  scale: 1
  num_threads_width: 10
  num_threads_height: 10

screen:
  width: 40
  height: 30

loader:
  Sandbox

image_output_dir: "../"

image_action:
  None

tone_map:
  Reinhard:
    default: true
""")
    scene = tmp_path / "scene.yml"
    scene.write_text("""---
textures:
  Grey:   # keys are lower-cased on load
    SolidColor:
        color:
          color: [ 0.5, 0.5, 0.5 ]
materials:
  grey:
    Lambertian:
      texture: grey
geometry:
  SphereA:
    Sphere:
      pos: [ 0, 1.0, 0 ]
      radius: 1
      material: grey
""")
    out = tmp_path / "flat.json"
    run(racer_render, "--config", str(cfg_file), "--scene", str(scene), "--dump-flat", str(out))
    d = json.loads(out.read_text())
    assert d["scene"]["object_keys"] == ["spherea"] and d["scene"]["prim_data"] == [0.0, 1.0, 0.0, 1.0, 0.0]
    assert d["render"] == [9, 7] and d["preview"] == [4, 3, 4] and d["tone_map_type"] == capi.RC_TONE_REINHARD
    assert (d["width"], d["height"]) == (40, 30)
    # camera defaults of CameraData::merge (camera.rs:404-464): vfov 20, pos 0, look_at -z
    assert d["camera"][0:3] == [0.0, 0.0, 0.0] and d["camera"][21] == 20.0

    bad = tmp_path / "bad.yml"
    bad.write_text(scene.read_text().replace("material: grey", "material: gold"))
    r = run(racer_render, "--config", str(cfg_file), "--scene", str(bad), "--dump-flat", str(out), check=False)
    assert r.returncode == 4 and "Unknown Material gold." in r.stderr                 # UnknownMaterial
    bad.write_text(scene.read_text().replace("texture: grey", "texture: blue"))
    r = run(racer_render, "--config", str(cfg_file), "--scene", str(bad), "--dump-flat", str(out), check=False)
    assert r.returncode == 9 and 'Failed to find texture "blue" for lambertian material "grey"' in r.stderr   # SceneLoad
    bad.write_text("textures: {}\nmaterials: {}\n")
    r = run(racer_render, "--config", str(cfg_file), "--scene", str(bad), "--dump-flat", str(out), check=False)
    assert r.returncode == 3 and "missing field `geometry`" in r.stderr               # Configuration
    r = run(racer_render, "--config", str(tmp_path / "nope.yml"), "--scene", str(scene), check=False)
    assert r.returncode == 3
    r = run(racer_render, "--bogus", check=False)
    assert r.returncode == 10                                                         # ArgumentParsingError


def test_host_library_exports_the_renderer_mirror(racer_render):
    """libracer_host.so carries the C++ mirror of the Renderer trait and friends."""
    out = subprocess.run(["nm", "-DC", os.path.join(PKG, "libracer_host.so")], capture_output=True, text=True).stdout
    for sym in ("racer::CudaRenderer::render", "racer::CudaPreviewRenderer::render", "racer::make_renderer",
                "racer::SceneData::load_yml", "racer::Config::from_file", "racer::make_camera", "racer::merge_camera",
                "racer::ScreenBuffer::update", "racer::save_png", "racer::sha256_hex_upper"):
        assert sym in out, sym
    # nothing of the oracle is linked into the host
    assert "oracle" not in subprocess.run(["ldd", os.path.join(PKG, "libracer_host.so")], capture_output=True, text=True).stdout


def read_png(path):
    data = open(path, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, kind = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert zlib.crc32(kind + body) == struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0]
        if kind == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 6)
        if kind == b"IDAT":
            idat += body
        pos += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 4 * w)
    assert not raw[:, 0].any()
    return raw[:, 1:].reshape(h, w, 4)


@pytest.mark.gpu
def test_racer_render_png_equals_the_harness_path(racer_render, tmp_path, cfg, renderer):
    """Same scene, seed and parameters through the C++ host (racer_render -> CudaRenderer::render ->
    ScreenBuffer -> SavePng) and through the Python harness: identical RGBA bytes; the default file name
    is the upper-case SHA-256 of those bytes (png.rs:36-41)."""
    import hashlib
    w, h, spp = 96, 64, 16
    out = tmp_path / "img.png"
    r = run(racer_render, "--config", CONFIG, "--scene", scene_path("cornell_box"), "--width", str(w), "--height", str(h),
            "--samples", str(spp), "--seed", "5", "--out", str(out))
    assert "It took" in r.stdout and "Saved image to" in r.stdout
    png = read_png(out)
    job = harness.prepare_job(scene_path("cornell_box"), cfg, w, h)
    renderer.upload(job)
    img = renderer.render(harness.make_params(w, h, spp, cfg.render.max_depth, seed=5, specialize=2))
    rgba, _ = renderer.postprocess(job.tone_map, img)
    assert np.array_equal(png, rgba)
    # SavePng naming + the preview renderer through the same binary
    cfg_file = tmp_path / "config.yml"
    cfg_file.write_text(open(CONFIG).read().replace("image_action: None", "image_action: SavePng")
                        .replace("image_output_dir: ../", f"image_output_dir: {tmp_path}"))
    r = run(racer_render, "--config", str(cfg_file), "--scene", scene_path("cornell_box"), "--width", "120", "--height", "80",
            "--preview", "--seed", "5")
    saved = r.stdout.strip().split("Saved image to: ")[1]
    body = read_png(saved)
    assert os.path.basename(saved) == hashlib.sha256(body.tobytes()).hexdigest().upper() + ".png"
    sw, sh = harness.preview_scales(cfg, 120, 80)
    job = harness.prepare_job(scene_path("cornell_box"), cfg, 120, 80)
    renderer.upload(job)
    pv = renderer.render_preview(harness.make_params(120, 80, cfg.preview.samples, cfg.preview.max_depth, seed=5), sw, sh)
    assert np.array_equal(body, renderer.postprocess(job.tone_map, pv)[0])


def _canon(v):
    """PyYAML value -> the shape --dump-yaml prints: lower-cased keys, scalars by numeric value or text."""
    if isinstance(v, dict):
        return {str(k).lower(): _canon(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_canon(x) for x in v]
    return _scalar(v)


def _scalar(v):
    if v is None:
        return None
    if isinstance(v, bool):
        return str(v).lower()
    if isinstance(v, (int, float)):
        return float(v)
    try:
        return float(v)
    except ValueError:
        return str(v)


def _canon_cpp(v):
    if isinstance(v, dict):
        return {k: _canon_cpp(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_canon_cpp(x) for x in v]
    return _scalar(v)


BLOCK_STYLE = """\
# block style, as the reference's own scene files are written (resources/scenes/*.yml)
materials:
  Ground:
    Lambertian:
      texture: ground   # trailing comment
  "quoted name":
    Metal:
      texture: 'single quoted'
      fuzz: 1.0e-1
geometry:
  floor:
    XzRect:
      x0: -1000
      x1: +1000.5
      z0: 0
      z1: 1
      k: 0
      material: Ground
  list_block:
    Sphere:
      pos:
        - 0
        - -1.5
        - 2
      radius: 0.5
      material: "quoted name"
  nested_flow: {Box: {min: {pos: [0, 0, 0]}, max: {pos: [1, 2, 3]}, material: Ground}}
empty_map: {}
empty_seq: []
flag: true
"""


@pytest.mark.parametrize("name", ["config.yml"] + ["scenes/" + s + ".yml" for s in
                                                     ("clown", "cornell_box", "emissive", "noise_and_textures", "sandbox_boxes",
                                                      "three_balls", "two_balls")] + ["<block>"])
def test_yaml_reader_agrees_with_pyyaml(racer_render, tmp_path, name):
    """yaml_lite.hpp (the C++ host's reader for config.yml / scene files, replacing the reference's serde_yaml
    use at src/config.rs:306 and src/scene/yml.rs:213) parses every fixture, and a block-style document in the
    style of the reference's own files, to the same tree as PyYAML."""
    yaml = pytest.importorskip("yaml")
    if name == "<block>":
        path = tmp_path / "block.yml"
        path.write_text(BLOCK_STYLE)
    else:
        path = os.path.join(ROOT, "tests", "golden", name)
    out = run(racer_render, "--dump-yaml", str(path)).stdout
    got = _canon_cpp(json.loads(out))
    want = _canon(yaml.safe_load(open(path).read()))
    assert got == want
