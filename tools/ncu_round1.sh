set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r01_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_bench.log 2>&1
$NCU --set full --import-source on -k regex:pp_kernel -s 3 -c 1 -o gpurun_out/r01_pp_rm_u python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_pp.log 2>&1
$NCU --set full --import-source on -k regex:panels_on_points_kernel -c 1 -o gpurun_out/r01_panels_on_points python tools/panel_bench.py > gpurun_out/ncu_pop.log 2>&1
$NCU --set full --import-source on -k regex:points_on_panels_kernel -c 1 -o gpurun_out/r01_points_on_panels python tools/panel_bench.py > gpurun_out/ncu_ptp.log 2>&1
$NCU --set full --import-source on -k regex:coeff_kernel -s 8 -c 1 -o gpurun_out/r01_coeff python tools/panel_bench.py > gpurun_out/ncu_coeff.log 2>&1
$NCU --set full --import-source on -k regex:clear_inner_kernel -s 2 -c 1 -o gpurun_out/r01_clear_inner python tools/panel_bench.py > gpurun_out/ncu_cil.log 2>&1
ls -la gpurun_out/*.ncu-rep
