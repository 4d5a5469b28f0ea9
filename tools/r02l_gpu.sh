#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r02l
mkdir -p $O
PY=$(which python)
( time timeout 1800 $PY -m pytest tests -m gpu -q ) > $O/pytest.log 2>&1
tail -6 $O/pytest.log
( time timeout 1200 $PY bench.py --steps 20 --warmup 5 ) > $O/bench.json 2> $O/bench.err
tail -c 1800 $O/bench.json; tail -2 $O/bench.err | cut -c1-300
M=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,smsp__sass_thread_inst_executed_op_fp64_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_integer_pred_on.sum,smsp__sass_thread_inst_executed_op_conversion_pred_on.sum,smsp__sass_thread_inst_executed_op_control_pred_on.sum,smsp__sass_thread_inst_executed_op_memory_pred_on.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed_pipe_fp64.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__registers_per_thread,smsp__warps_active.avg.per_cycle_active,launch__grid_size,launch__block_size
for a in exact contract; do
  timeout 600 ncu --metrics $M --clock-control none -k regex:sdf_column_kernel -c 1 --csv --log-file $O/k1_ops_$a.csv $PY tools/profile_dense.py --grid 1000 --repeat 1 --arithmetic $a > $O/k1_ops_$a.log 2>&1
done
timeout 900 ncu --metrics $M --clock-control none -k regex:tile_color_kernel -c 1 --csv --log-file $O/k6_ops.csv $PY tools/ab_time.py --grid 500 --texture 4096 > $O/k6_ops.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tile_color_kernel -c 1 -o $O/k6_full $PY tools/ab_time.py --grid 500 --texture 4096 > $O/k6_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:sdf_column_kernel -c 1 -o $O/k1_exact_full $PY tools/profile_dense.py --grid 1000 --repeat 1 --arithmetic exact > $O/k1_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:sdf_column_kernel -c 1 -o $O/k1_contract_full $PY tools/profile_dense.py --grid 1000 --repeat 1 --arithmetic contract >> $O/k1_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench.csv $PY bench.py --steps 2 --warmup 1 --batch 4 > $O/bench_under_ncu.log 2>&1
( UPG_TRACE=1 timeout 300 $PY tools/png_probe.py ) > $O/png_probe.json 2> $O/png_probe.err; cat $O/png_probe.json | cut -c1-900; grep "png" $O/png_probe.err | tail -6
( timeout 1200 $PY tools/baseline_table.py --out $O/baseline_table.json --configs 1,2,3 ) > $O/baseline_table.log 2>&1; tail -3 $O/baseline_table.log | cut -c1-1500
ls -la $O
