#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02f
mkdir -p $O
PY=$(which python)
( time timeout 1500 $PY -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -4 $O/pytest.log
( timeout 600 $PY tools/flake_probe.py --res 150 --repeat 150 --planets torus:craters,doubletorus:islands ) > $O/flake_default.log 2>&1
echo "default (early water on): $(grep -c MISMATCH $O/flake_default.log) mismatches in 600 planets"
( time timeout 1200 $PY bench.py --steps 20 --warmup 5 ) > $O/bench.json 2> $O/bench.err
tail -c 3000 $O/bench.json
( time timeout 600 $PY bench.py --impl reference --steps 3 --warmup 1 ) > $O/bench_ref.json 2> $O/bench_ref.err
tail -c 600 $O/bench_ref.json; tail -3 $O/bench_ref.err
