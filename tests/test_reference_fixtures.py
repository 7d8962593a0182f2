"""Oracle against fixtures produced by the UNMODIFIED reference (tools/ref_fixtures.mjs under Node >= 18 with gl-matrix 3.4.4).

The build container has no JS engine, so tests/golden/ref/ is normally absent and these tests skip with that reason — the
oracle then stays "parity unpinned" (DESIGN.md §6).  On the first box where `npx tsx tools/ref_fixtures.mjs --reference <clone>`
has been run, they compare every case of tests/golden/manifest.json bit for bit and settle the one open question of SURVEY.md
Appendix B (whether gl-matrix's vec3.length is Math.hypot or a plain sqrt) from the recorded function source."""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "golden", "ref")
META = os.path.join(REF, "ref_meta.json")
ROOT = os.path.dirname(HERE)


def _cases():
    with open(os.path.join(HERE, "golden", "manifest.json")) as fh:
        return json.load(fh)["cases"]


def test_fixture_recipe_is_committed_and_covers_the_manifest():
    src = open(os.path.join(ROOT, "tools", "ref_fixtures.mjs")).read()
    assert "src/workers/raymarchWorker.ts" in src and "manifest.json" in src and "vec3_length_uses_hypot" in src
    assert "self.onmessage(" in src, "the recipe must drive the reference's own worker body, not a re-implementation"
    for name in ("binding.gyp", "package.json"):
        assert os.path.exists(os.path.join(ROOT, "addon", name))
    pkg = json.load(open(os.path.join(ROOT, "addon", "package.json")))
    assert pkg["dependencies"]["gl-matrix"] == "3.4.4"  # the reference's pin (package-lock.json:1258-1263)


@pytest.mark.skipif(not os.path.exists(META), reason="no reference-generated fixtures (tests/golden/ref): Node.js is not available in this image; "
                    "run tools/ref_fixtures.mjs on a box with Node >= 18 to pin the oracle")
@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_oracle_equals_reference_fixture(oracle, case):
    meta = json.load(open(META))
    oracle.lib().orc_set_length_mode(1 if meta["vec3_length_uses_hypot"] else 0)
    try:
        s = oracle.OracleScene()
        if case["synthetic"]:
            s.load_synthetic(*case["synthetic"])
        else:
            s.load_preset(case["preset"])
        s.build_accel(case["accel"]).set_camera(case["pitch"], case["yaw"]).set_time(case["time"])
        f = s.render(case["W"], case["H"], case["alg"], step_size=case["step"], overshoot=case["over"])
        raw = np.fromfile(os.path.join(REF, case["name"] + ".refbin"), np.uint8)
        n = case["W"] * case["H"]
        assert raw.size == 8 * n
        assert np.array_equal(raw[:n], f.depth), "depth bytes differ from the reference"
        assert np.array_equal(raw[n:4 * n], f.normal), "normal bytes differ from the reference"
        assert np.array_equal(raw[4 * n:6 * n].view(np.uint16), f.sdfEval), "SDF-call counters differ from the reference"
        assert np.array_equal(raw[6 * n:].view(np.uint16), f.iters), "iteration counters differ from the reference"
    finally:
        oracle.lib().orc_set_length_mode(1)
