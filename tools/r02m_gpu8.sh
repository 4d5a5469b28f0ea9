#!/bin/bash
# 8-GPU call: multi-device tests at 2 / 3 / 4 / 8 devices, the multi-device planet 1 -> 8, bench under torchrun at N = 8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02m
mkdir -p $O
PY=$(which python)
nvidia-smi -L > $O/gpus.txt
( time timeout 900 $PY -m pytest tests/test_gpu_multi.py -m gpu -q ) > $O/pytest_multi.log 2>&1
tail -6 $O/pytest_multi.log
( timeout 900 $PY tools/multi_probe.py --repeat 3 ) > $O/multi_probe.log 2> $O/multi_probe.err
tail -1 $O/multi_probe.log | cut -c1-1500
( UPG_TRACE=1 timeout 300 $PY tools/multi_probe.py --repeat 2 --counts 8 ) > $O/multi_probe_trace.log 2> $O/multi_probe_trace.err
grep "multi-device" $O/multi_probe_trace.err | tail -8
( time timeout 1500 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 ) > $O/bench_n8.json 2> $O/bench_n8.err
tail -c 2500 $O/bench_n8.json; tail -3 $O/bench_n8.err | cut -c1-300
( timeout 900 $PY tools/baseline_table.py --out $O/baseline_table_c5.json --configs 5 ) > $O/baseline_c5.log 2>&1; tail -1 $O/baseline_c5.log | cut -c1-1200
