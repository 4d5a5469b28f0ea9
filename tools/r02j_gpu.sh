#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02j
mkdir -p $O
PY=$(which python)
( time timeout 1800 $PY -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -8 $O/pytest.log
( time timeout 1200 $PY bench.py --steps 20 --warmup 5 ) > $O/bench.json 2> $O/bench.err
tail -c 2500 $O/bench.json; tail -3 $O/bench.err | cut -c1-300
