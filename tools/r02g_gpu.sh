#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02g
mkdir -p $O
PY=$(which python)
nvidia-smi -L > $O/gpus.txt; nvidia-smi topo -m > $O/topo.txt 2>&1
( time timeout 900 $PY -m pytest tests/test_gpu_multi.py -m gpu -q -x ) > $O/pytest_multi.log 2>&1
tail -15 $O/pytest_multi.log
( UPG_TRACE=1 timeout 900 $PY tools/multi_probe.py --repeat 3 ) > $O/multi_probe.log 2> $O/multi_probe.err
grep -v "^\[upg\]" $O/multi_probe.log | tail -5; grep "multi-device" $O/multi_probe.err | tail -12
