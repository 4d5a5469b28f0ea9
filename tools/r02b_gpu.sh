#!/bin/bash
# round 2, GPU call B: the flake hunt (no-host-sync ordering under poison / sanitizers) + the remaining GPU tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02b
mkdir -p $O
PY=$(which python)
( UPG_EARLY_WATER=1 timeout 600 $PY tools/flake_probe.py --res 150 --repeat 30 ) > $O/early_water.log 2>&1
( UPG_EARLY_WATER=1 UPG_POISON=ff timeout 600 $PY tools/flake_probe.py --res 150 --repeat 30 ) > $O/early_water_poison_ff.log 2>&1
( UPG_EARLY_WATER=1 UPG_POISON=5a timeout 600 $PY tools/flake_probe.py --res 150 --repeat 30 ) > $O/early_water_poison_5a.log 2>&1
( UPG_POISON=ff timeout 600 $PY tools/flake_probe.py --res 150 --repeat 10 ) > $O/default_poison_ff.log 2>&1
grep -c MISMATCH $O/*.log
( timeout 900 compute-sanitizer --tool initcheck --track-unused-memory no $PY tools/flake_probe.py --res 64 --planets sphere:lakes,torus:craters ) > $O/initcheck.log 2>&1
tail -5 $O/initcheck.log
( UPG_EARLY_WATER=1 timeout 900 compute-sanitizer --tool memcheck $PY tools/flake_probe.py --res 64 --planets sphere:lakes,torus:craters --repeat 3 ) > $O/memcheck_early.log 2>&1
tail -3 $O/memcheck_early.log
( time timeout 1500 $PY -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -15 $O/pytest.log
