"""The pin hook for the absent FastNoiseLite header (tools/fastnoise_pin.py, `make -C oracle ref FASTNOISE=<header>`), exercised with a
FIXTURE header that has the real one's interface (tests/golden/mock_fastnoise/FastNoiseLite.h): the reference's translation units
build against it, every generator is compared through both libraries, the private constant table is extracted."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/sources"), reason="builds the reference's sources: needs the checkout")
def test_pin_hook_builds_compares_and_extracts(tmp_path):
    header = os.path.join(ROOT, "tests", "golden", "mock_fastnoise", "FastNoiseLite.h")
    golden = str(tmp_path / "randvecs.npy")
    report = str(tmp_path / "report.json")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fastnoise_pin.py"), "--header", header, "--golden", golden, "--json", report], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    out = json.load(open(report))
    assert out["differing_generators"] == [] and len(out["rows"]) > 80
    assert all(row["bit_identical"] == 1.0 for row in out["rows"])  # the fixture delegates to the restatement
    assert out["randvecs3d_equal"] is False  # ... but exposes its own table: the tool read it from the "upstream" header
    t = np.load(golden)
    assert t.shape == (256, 4) and np.allclose(t[:, 0], np.cos(np.arange(256)), atol=1e-6) and (t[:, 3] == 0).all()
