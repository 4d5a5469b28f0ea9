#!/bin/bash
# round 2, GPU call A: sanity of the new allocator / status words / contract build, full-size parity, flake hunt, K1 op counts
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi -L > $O/gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $O/pytest.log 2>&1
tail -5 $O/pytest.log
timeout 600 python tools/ab_k1.py --grid 1000 > $O/ab_k1_sphere_lakes_1000.json 2> $O/ab_k1.err
timeout 300 python tools/ab_k1.py --grid 500 --shape torus --elevation craters > $O/ab_k1_torus_craters_500.json 2>> $O/ab_k1.err
timeout 300 python tools/ab_k1.py --grid 500 --shape sphere --elevation simple > $O/ab_k1_sphere_simple_500.json 2>> $O/ab_k1.err
cat $O/ab_k1_sphere_lakes_1000.json
# the experimental ordering of round 1 (water extraction beside the band kernel, no host sync before the workers), 40 rounds
( UPG_EARLY_WATER=1 timeout 900 python tools/flake_probe.py --res 150 --repeat 40 ) > $O/early_water.log 2>&1
( timeout 600 python tools/flake_probe.py --res 150 --repeat 20 ) > $O/default_order.log 2>&1
tail -2 $O/early_water.log $O/default_order.log
( timeout 900 compute-sanitizer --tool initcheck --track-unused-memory no python tools/flake_probe.py --res 64 --planets sphere:lakes,torus:craters ) > $O/initcheck.log 2>&1
tail -5 $O/initcheck.log
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,smsp__sass_thread_inst_executed_op_integer_pred_on.sum,smsp__sass_thread_inst_executed_op_conversion_pred_on.sum,smsp__sass_thread_inst_executed_op_control_pred_on.sum,smsp__sass_thread_inst_executed_op_memory_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,smsp__warps_active.avg.per_cycle_active
for a in exact contract; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:sdf_column_kernel -c 1 --csv --log-file $O/k1_ops_$a.csv python tools/profile_dense.py --grid 1000 --repeat 1 --arithmetic $a > $O/k1_ops_$a.log 2>&1
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:sdf_column_kernel -c 1 -o $O/k1_contract_full python tools/profile_dense.py --grid 1000 --repeat 1 --arithmetic contract > $O/k1_full.log 2>&1
ls -la $O
