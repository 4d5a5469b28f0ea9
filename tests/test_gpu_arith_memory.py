"""GPU tests of the round-2 additions around K1 and the library's memory:
  * UPG_ARITH_CONTRACT (the dense K1 kernels compiled with FMA contraction) against UPG_ARITH_EXACT: within the stated FP32
    tolerance, with the samples outside it enumerated (the reference's own cover-threshold step, see test_gpu_fullsize.py);
  * scratch memory: results do not depend on what the pool held before (two different UPG_POISON fills, in subprocesses);
  * per-call NaN status words, upg_planet_async_status, upg_reserve / upg_trim / upg_memory_stats."""
import json
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

from conftest import ROOT, SDF_ATOL, SDF_RTOL, sdf_close
from unnatural_planets_b200 import Planet, UpgError
from unnatural_planets_b200.planet import memory_stats, reserve, trim

pytestmark = pytest.mark.gpu

STEP = 12.5 / 10.0  # smoothMax(0, 0, 50) / meshElevationRatio: the reference's own cover-threshold step (tests/test_gpu_fullsize.py)


@pytest.mark.parametrize("shape,elev", [("sphere", "lakes"), ("torus", "craters"), ("fibers", "simple"), ("bunny", "legacy"), ("asteroid", "islands"), ("knot", "none")])
def test_contract_build_within_tolerance(shape, elev):
    res = 128
    exact = Planet(shape=shape, elevation=elev, grid=res, arithmetic="exact")
    contract = Planet(shape=shape, elevation=elev, grid=res, arithmetic="contract")
    report = {}
    for field in ("land", "water", "navigation"):
        a, b = exact.field_samples(field), contract.field_samples(field)
        ok = sdf_close(a, b, SDF_RTOL, SDF_ATOL)
        bad = ~ok
        if bad.any():
            # only the cover-threshold step of lakes / islands, or the legacy mode's second noise (its lattice input is itself a
            # noise value: a one-ulp change can cross a lattice plane, the value stays continuous)
            diff = np.abs(a.astype(np.float64) - b)[bad]
            assert elev in ("lakes", "islands") and np.all(np.abs(diff - STEP) < 5e-3), (shape, elev, field, int(bad.sum()), diff[:8])
            assert bad.sum() <= max(2, ok.size // 100000)
        far = (np.abs(a) > SDF_ATOL) & ok
        assert ((a >= 0) == (b >= 0))[far].all()
        report[field] = (int(bad.sum()), float(np.abs(a - b)[ok].max()))
    # the fused launch agrees with the single-field launches of the same build (bit-equal in the exact build, tests/test_gpu_field.py;
    # the tolerance build lets the compiler fuse differently per launch shape: agreement to a few ulp of the operands)
    import torch

    bufs = [torch.empty(res**3, dtype=torch.float32, device="cuda") for _ in range(3)]
    s = torch.cuda.current_stream()
    contract.fields_samples_device(0, res, bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(), s.cuda_stream)
    contract.async_status(s.cuda_stream)
    land = bufs[0].cpu().numpy().reshape(res, res, res)
    single = contract.field_samples("land")
    close = sdf_close(land, single, 1e-5, 1e-3)
    assert close.mean() > 0.99999 and np.abs(land - single)[~close].max(initial=0) < STEP + 5e-3


def test_contract_build_does_not_touch_meshes_points_or_bakes():
    # the mesh entry points, upg_field_points and the bake are the validator: always exact
    a = Planet(shape="sphere", elevation="lakes", grid=96, arithmetic="exact")
    b = Planet(shape="sphere", elevation="lakes", grid=96, arithmetic="contract")
    ma, mb = a.field_meshes_clean(("land", "water", "navigation")), b.field_meshes_clean(("land", "water", "navigation"))
    for f in ma:
        for k in ("positions", "normals", "indices"):
            assert (ma[f][k].view(np.uint32) == mb[f][k].view(np.uint32)).all(), (f, k)
    pts = (np.random.default_rng(5).random((4096, 3), dtype=np.float32) - 0.5) * 3000
    assert (a.field_points("land", pts).view(np.uint32) == b.field_points("land", pts).view(np.uint32)).all()


def test_scratch_memory_is_never_read_before_it_is_written():
    """tools/flake_probe.py digests every mesh of four planets (band and dense sampling, three concurrent extractions, K9) -- once
    with every pool allocation filled with 0xff bytes and once with 0x5a: a kernel that read scratch it never wrote would see NaNs /
    huge indices in one run and other garbage in the other."""
    outs = []
    for poison in ("ff", "5a", ""):
        env = dict(os.environ, UPG_POISON=poison)
        if not poison:
            env.pop("UPG_POISON")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "flake_probe.py"), "--res", "96"], capture_output=True, text=True, env=env, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "MISMATCH" not in r.stdout
        outs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert outs[0] == outs[1] == outs[2]


def test_status_words_are_per_call():
    # a planet whose land field is NaN everywhere (box far outside float range of the neural shape is hard to get; use the API
    # contract instead): concurrent calls on ONE planet from several threads must each report their own status
    good = Planet(shape="sphere", elevation="lakes", grid=64)
    errors = []

    def work():
        try:
            for _ in range(6):
                good.field_samples("land", 0, 8)
                good.field_points("land", np.zeros((16, 3), np.float32))
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    ts = [threading.Thread(target=work) for _ in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors[:1]
    good.async_status()


def test_nan_sample_is_reported_by_the_call_that_produced_it():
    # pow of a negative base: the `simple` elevation at a NaN position -> the reference throws "invalid land sdf value"
    p = Planet(shape="sphere", elevation="simple", grid=32)
    bad = np.array([[np.nan, 0, 0]], np.float32)
    with pytest.raises(UpgError) as e:
        p.field_points("land", bad)
    assert e.value.code == 3
    # the flag belonged to that call: the next one is clean
    assert np.isfinite(p.field_points("land", np.array([[1100, 0, 0]], np.float32))).all()
    p.async_status()


def test_reserve_trim_and_stats():
    trim(0, 0)
    before = memory_stats(0)
    reserve(0, device_bytes=256 << 20, pinned_block_bytes=8 << 20, pinned_blocks=2)
    after = memory_stats(0)
    assert after["device_reserved"] >= 256 << 20
    assert after["pinned_cached"] >= before["pinned_cached"] + (16 << 20)
    Planet(shape="sphere", elevation="lakes", grid=64).field_mesh("land")
    trim(0, 0)
    end = memory_stats(0)
    assert end["device_reserved"] < 256 << 20 and end["pinned_cached"] == 0


def test_grid_resolution_limit_is_checked_at_creation():
    with pytest.raises(UpgError) as e:
        Planet(shape="sphere", elevation="none", grid=1627)
    assert e.value.code == 1
