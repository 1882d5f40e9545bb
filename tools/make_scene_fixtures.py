#!/usr/bin/env python
"""Regenerates tests/golden/scenes/*.yml, tests/golden/config.yml and the image
asset from the reference checkout (run in the build container only; the GPU box
has no /root/reference).  The scene/config files are DATA in the reference's
YAML formats — the inputs every parity test and the bench render — re-emitted
through a YAML round trip (comments and layout are not preserved, values are).
"""
import os
import shutil
import sys

import yaml

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

for name in sorted(os.listdir(os.path.join(REF, "resources", "scenes"))):
    with open(os.path.join(REF, "resources", "scenes", name)) as f:
        doc = yaml.safe_load(f)
    with open(os.path.join(OUT, "scenes", name), "w") as f:
        f.write(f"# scene data of racer-tracer resources/scenes/{name}, re-emitted by tools/make_scene_fixtures.py\n")
        yaml.safe_dump(doc, f, sort_keys=False, default_flow_style=None, width=100)
with open(os.path.join(REF, "racer-tracer", "config.yml")) as f:
    doc = yaml.safe_load(f)
with open(os.path.join(OUT, "config.yml"), "w") as f:
    f.write("# values of racer-tracer/config.yml, re-emitted by tools/make_scene_fixtures.py\n")
    yaml.safe_dump(doc, f, sort_keys=False, default_flow_style=None, width=100)
shutil.copyfile(os.path.join(REF, "resources", "images", "earthmap.jpg"),
                os.path.join(OUT, "resources", "images", "earthmap.jpg"))
os.chmod(os.path.join(OUT, "resources", "images", "earthmap.jpg"), 0o644)
print("fixtures written to", os.path.normpath(OUT))
