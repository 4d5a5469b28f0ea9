#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02k
mkdir -p $O
PY=$(which python)
( time timeout 900 $PY -m pytest tests/test_gpu_png.py tests/test_shim.py -m gpu -q -x ) > $O/pytest_png_shim.log 2>&1
tail -12 $O/pytest_png_shim.log
timeout 600 $PY tools/png_probe.py > $O/png_probe.json 2> $O/png_probe.err; cat $O/png_probe.json; tail -3 $O/png_probe.err
