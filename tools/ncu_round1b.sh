set -x
python bench.py > gpurun_out/r01_bench_c3.json 2> gpurun_out/r01_bench_c3.err
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r01_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_bench.log 2>&1
$NCU --set full --import-source on -k regex:pp_kernel -s 3 -c 1 -o gpurun_out/r01_pp_rm_ug python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-variants > gpurun_out/ncu_pp_ug.log 2>&1
python tools/path_table.py --out gpurun_out/r01_path_table.md > gpurun_out/path_table.log 2>&1
tail -c 300 gpurun_out/r01_bench_c3.json
