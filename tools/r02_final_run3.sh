# last evidence pass of round 2 (one B200): tests, bench line, path table, substep table, panel kernels
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/r02_bench_c3.json 2> gpurun_out/r02_bench_c3.err; tail -c 200 gpurun_out/r02_bench_c3.err
python tools/path_table.py --out gpurun_out/r02_path_table.md > gpurun_out/r02_path_table.log 2>&1; tail -1 gpurun_out/r02_path_table.log
python tools/c2_substep.py --out gpurun_out/r02_c2_substep.md > /dev/null 2>&1
python tools/panel_bench.py > gpurun_out/r02_panel_bench.txt 2>&1
build/flow_over_circle_step 10000 50; build/speedtest20kx10 | tail -3
oracle/_ref/ref_dropin_v2 | tail -1; oracle/_ref/ref_external_seam_v2 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()"
