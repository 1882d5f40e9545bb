import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
SCENES = ["three_balls", "emissive", "noise_and_textures", "cornell_box", "clown"]   # the BASELINE configs
# + the reference's Sandbox scene (Box / RotateY / Translate instances) in YAML form
# + the reference's Random scene (scene/random.rs: ~480 spheres, moving spheres + ray time, lens), §8(f)-4
# + the reference's Sandbox LOADER (scene/sandbox.rs: cornell_box.yml + two rotated, translated boxes), the scene
#   racer-tracer/config.yml selects by default
SCENES_X = SCENES + ["sandbox_boxes", "random", "sandbox"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def scene_path(name):
    if name == "random":      # SceneLoaderConfig::Random: generated, not a file
        return name
    if name == "sandbox":     # SceneLoaderConfig::Sandbox: cornell_box.yml + programmatic additions
        return "sandbox:" + os.path.join(GOLDEN, "scenes", "cornell_box.yml")
    return os.path.join(GOLDEN, "scenes", name + ".yml")


@pytest.fixture(scope="session")
def cfg():
    from racer_tracer_b200 import harness
    return harness.load_config(os.path.join(GOLDEN, "config.yml"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def cuda_lib():
    """libracer_cuda.so, built in-tree if missing (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge
    from racer_tracer_b200 import capi
    if not os.path.exists(capi.LIB_PATH):
        ge.build()
    return capi.load()


@pytest.fixture(scope="session")
def renderer(cuda_lib):
    from racer_tracer_b200 import harness
    r = harness.CudaRenderer([0])
    yield r
    r.close()
