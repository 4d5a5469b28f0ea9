#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02i
mkdir -p $O
PY=$(which python)
for v in "" mb4 mb6; do
  if [ -z "$v" ]; then lib=""; else lib="$PWD/build_variants/libupg_$v.so"; fi
  UPG_LIBRARY=$lib timeout 600 $PY tools/ab_k1.py --grid 1000 > $O/ab_k1_${v:-default}.json 2>> $O/ab_k1.err
  cut -c1-160 $O/ab_k1_${v:-default}.json
done
( time timeout 1800 $PY -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -6 $O/pytest.log
