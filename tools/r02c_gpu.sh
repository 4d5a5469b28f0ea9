#!/bin/bash
# round 2, GPU call C: which ingredient of the no-host-sync ordering loses a vertex?
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02c
mkdir -p $O
PY=$(which python)
run() { # name, env...
  name=$1; shift
  ( env "$@" UPG_EARLY_WATER=1 timeout 400 $PY tools/flake_probe.py --res 150 --repeat 100 --planets torus:craters,doubletorus:islands ) > $O/$name.log 2>&1
  echo "$name: $(grep -c MISMATCH $O/$name.log) mismatches"
}
run base UPG_X=0
run base_again UPG_X=0
run no_event_reuse UPG_POOL_REUSE_OFF=1
run no_opportunistic UPG_POOL_REUSE_OFF=2
run no_internal UPG_POOL_REUSE_OFF=4
run no_reuse_at_all UPG_POOL_REUSE_OFF=7
run serial_workers UPG_DEBUG_WORKERS=1
run worker_host_wait UPG_DEBUG_WORKERS=2
( timeout 400 $PY tools/flake_probe.py --res 150 --repeat 100 --planets torus:craters,doubletorus:islands ) > $O/default_order.log 2>&1
echo "default_order: $(grep -c MISMATCH $O/default_order.log) mismatches"
