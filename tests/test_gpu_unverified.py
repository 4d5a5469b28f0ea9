"""GPU checks of code written AFTER the round's GPU budget was spent (round 2: the 8-GPU run consumed what was left, see
profiles/r02_m_multi_gpu8.md).  They have never run on a B200; the driver's round-end run is their first.  Every file of
tests/unverified_gpu/ therefore runs in its OWN python process (a kernel fault there cannot take the CUDA context of the verified tests
with it) under a timeout, and is reported non-strictly: XPASS = every check of the file holds on hardware, XFAIL = one does not (or the file
could not run) -- neither hides or stops the verified tests behind `-x`.  DESIGN.md section 9 ("not verified on hardware") lists what
these cover."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKS = ["check_mesh_chunking.py", "check_bake_on_xatlas_chunks.py", "check_bake_export_png.py", "check_multi_host_exchange.py"]


@pytest.mark.parametrize("name", CHECKS)
@pytest.mark.xfail(strict=False, reason="written after the round's GPU budget was spent: first run on a B200 (XPASS = holds on hardware)")
def test_first_run_on_hardware(name):
    path = os.path.join("tests", "unverified_gpu", name)
    r = subprocess.run([sys.executable, "-m", "pytest", path, "-q", "-m", "gpu", "-p", "no:cacheprovider", "-rA"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    sys.stdout.write(r.stdout[-6000:])
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0, r.stdout[-3000:]
    if " passed" not in r.stdout:
        pytest.skip("nothing ran (%s)" % (r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "no output"))
