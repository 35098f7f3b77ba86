"""Writes tests/golden/cnidaria_panel.json: the primer pairs of the reference's cnidaria panel
(/root/reference/panels/cnidaria.yaml) as the PCRParams the reference derives from them
(src/pcr/preconfigured.rs:216-334 through sharkmer_b200.panels.load_panel_file).  The GPU box has
no /root/reference, so the C4 test (tests/test_gpu_c4.py) reads this fixture instead.

    python tests/golden/make_panel_fixture.py
"""
import dataclasses
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from sharkmer_b200 import panels  # noqa: E402

SRC = "/root/reference/panels/cnidaria.yaml"

if __name__ == "__main__":
    prm = panels.load_panel_file(SRC)
    out = {"source": "panels/cnidaria.yaml", "primers": [dataclasses.asdict(p) for p in prm]}
    path = os.path.join(ROOT, "tests", "golden", "cnidaria_panel.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(path, len(prm), "primer pairs")
