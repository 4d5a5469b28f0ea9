#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02e
mkdir -p $O
PY=$(which python)
# the regression test against the round-1 flatten (expected to FAIL) and against the fix (expected to pass)
( UPG_LIBRARY=$PWD/build_variants/libupg_oldflatten.so timeout 600 $PY -m pytest tests/test_gpu_meshfix.py -m gpu -q -k flatten ) > $O/regression_old_flatten.log 2>&1
tail -3 $O/regression_old_flatten.log
( timeout 900 $PY tools/flake_probe.py --res 150 --repeat 400 --planets torus:craters --dump $O/dump ) > $O/fixed_default_order.log 2>&1
echo "fixed default order: $(grep -c MISMATCH $O/fixed_default_order.log) mismatches in 800 planets"
( UPG_EARLY_WATER=1 timeout 900 $PY tools/flake_probe.py --res 150 --repeat 300 --planets torus:craters,doubletorus:islands ) > $O/fixed_early_water.log 2>&1
echo "fixed early water: $(grep -c MISMATCH $O/fixed_early_water.log) mismatches in 1200 planets"
( UPG_LIBRARY=$PWD/build_variants/libupg_oldflatten.so timeout 900 $PY tools/flake_probe.py --res 150 --repeat 400 --planets torus:craters ) > $O/old_default_order.log 2>&1
echo "old flatten, default order: $(grep -c MISMATCH $O/old_default_order.log) mismatches in 800 planets"
( time timeout 1500 $PY -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -6 $O/pytest.log
UPG_TRACE=1 timeout 300 $PY tools/mesh_e2e_probe.py > $O/mesh_e2e.log 2>&1; tail -12 $O/mesh_e2e.log
