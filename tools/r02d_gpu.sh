#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02d
mkdir -p $O
PY=$(which python)
( timeout 900 $PY tools/flake_probe.py --res 150 --repeat 400 --planets torus:craters --dump $O/dump ) > $O/default_order.log 2>&1
echo "default_order: $(grep -c MISMATCH $O/default_order.log) mismatches"; ls $O/dump 2>/dev/null | head
