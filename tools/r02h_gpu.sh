#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02h
mkdir -p $O
PY=$(which python)
( time timeout 1500 $PY -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
timeout 600 $PY tools/ab_k1.py --grid 1000 > $O/ab_k1_sphere_lakes_1000.json 2> $O/ab_k1.err; cat $O/ab_k1_sphere_lakes_1000.json | cut -c1-700
timeout 300 $PY tools/ab_k1.py --grid 500 --shape asteroid --elevation islands > $O/ab_k1_asteroid_islands_500.json 2>> $O/ab_k1.err; cut -c1-400 $O/ab_k1_asteroid_islands_500.json
