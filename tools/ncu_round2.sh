#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE B200).  Per dominant kernel: one `ncu --set full` capture of a single
# launch at bench size, with the hardware-counted FP32 op counters SURVEY 8(d) names added to the set; plus the launch
# list of the default bench command.  Reports come back in gpurun_out/; tools/ncu_summary.py turns them into profiles/*.csv.
set -x
EXTRA="smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,smsp__sass_thread_inst_executed_op_fp32_pred_on.sum,smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_fmaheavy.sum,smsp__inst_executed_pipe_fmalite.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,sm__cycles_elapsed.avg,sm__cycles_active.avg"
NCU="ncu --set full --metrics $EXTRA --clock-control none --import-source on"
cap() {  # name kernel-regex target...
  local name=$1 k=$2; shift 2
  $NCU -k regex:$k -s 1 -c 1 -f -o gpurun_out/r02_$name python tools/ncu_targets.py "$@" > gpurun_out/ncu_$name.log 2>&1
}
cap pp_rm_ug pp_kernel pp:rm:2:0
cap pp_rm_u pp_kernel pp:rm:1:0
cap pp_rm_u_uniform pp_kernel pp:rm:1:1
cap pp_v2_u_uniform pp_kernel pp:v2:1:1
cap pp_wl_u pp_kernel pp:wl:1:0
cap pp_exp_u pp_kernel pp:exp:1:0
cap pp_v3_u pp_kernel pp:v3:1:0
cap panels_on_points panels_on_points_kernel panels
cap points_on_panels points_on_panels_kernel ptpan
cap clear_inner clear_inner_kernel cil
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --no-strong > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/r02_*.ncu-rep
