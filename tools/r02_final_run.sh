# Round-2 evidence pass on ONE B200 (run under gpurun from the repo root; outputs land in gpurun_out/, the summaries
# the docs cite are copied from there into profiles/).  The per-core ncu captures of the pair-sum kernel are
# tools/ncu_round2.sh; the multi-GPU sweeps are tools/r02_sweeps.sh.
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -c 200 gpurun_out/r02_bench_c3.err
python tools/path_table.py --out gpurun_out/r02_path_table.md > gpurun_out/r02_path_table.log 2>&1; tail -1 gpurun_out/r02_path_table.log
python tools/c2_substep.py --out gpurun_out/r02_c2_substep.md > /dev/null 2>&1
python tools/panel_bench.py > gpurun_out/r02_panel_bench.txt 2>&1
python tools/core_sweep.py 1000000 2 > gpurun_out/r02_core_sweep.txt 2>&1
build/flow_over_circle_step 10000 50; build/speedtest20kx10 | tail -3
oracle/_ref/ref_dropin_v2 | tail -1; oracle/_ref/ref_external_seam_v2 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()"
bash tools/sanitize_round2.sh > /dev/null 2>&1; grep -c "0 errors\|0 hazards" gpurun_out/r02_sanitizer.txt
EXTRA="smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,sm__cycles_elapsed.avg,sm__cycles_active.avg,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"
NCU="ncu --set full --metrics $EXTRA --clock-control none --import-source on"
$NCU -k regex:panels_on_points_kernel -s 1 -c 1 -f -o gpurun_out/r02_panels_on_points python tools/ncu_targets.py panels > gpurun_out/ncu_pan.log 2>&1
$NCU -k regex:points_on_panels_kernel -s 1 -c 1 -f -o gpurun_out/r02_points_on_panels python tools/ncu_targets.py ptpan > gpurun_out/ncu_ptpan.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants --no-strong > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
