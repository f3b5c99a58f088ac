EXTRA="smsp__inst_executed_pipe_fma.sum,smsp__inst_executed_pipe_xu.sum,smsp__inst_executed_pipe_lsu.sum,sm__cycles_elapsed.avg,sm__cycles_active.avg"
ncu --set full --metrics $EXTRA --clock-control none --import-source on -k regex:coeff_kernel -s 1 -c 1 -f -o gpurun_out/r02_coeff python tools/ncu_targets.py coeff > gpurun_out/ncu_coeff.log 2>&1
tail -2 gpurun_out/ncu_coeff.log; ls -la gpurun_out/r02_coeff.ncu-rep
