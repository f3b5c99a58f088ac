#!/bin/bash
# Round-2 multi-GPU sweeps run by the builder (one gpurun call per GPU count):  tools/r02_sweeps.sh N
# C3 weak + C4 strong come from the default bench line; C5 strong (16.8M WL) and C5 weak (N_1 = 5 931 642, N_P = N_1 sqrt(P)).
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
[ "$N" = 1 ] && TR="python"
OPT="--steps 1 --warmup 1 --no-e2e --no-variants --no-strong"
if [ "$N" != 1 ] && [ "$N" != 8 ] && [ "$N" != 2 ]; then
  $TR ${N:+--master-port 29521} bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
fi
if [ "$N" != 1 ]; then
  $TR --master-port 29524 tools/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_n$N.txt 2>&1; tail -4 gpurun_out/r02_multi_gpu_check_n$N.txt
  $TR --master-port 29522 bench.py --gpus $N --workload c5 --scaling strong $OPT > gpurun_out/r02_c5_strong_n$N.json 2> gpurun_out/r02_c5_strong_n$N.err
  $TR --master-port 29523 bench.py --gpus $N --workload c5 --particles 5931642 --scaling weak $OPT > gpurun_out/r02_c5_weak_n$N.json 2> gpurun_out/r02_c5_weak_n$N.err
else
  python bench.py --workload c5 --particles 5931642 --scaling weak $OPT > gpurun_out/r02_c5_weak_n1.json 2> gpurun_out/r02_c5_weak_n1.err
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_*_n$N.json")):
    try:
        d = json.load(open(f))
        print(f, "value %.4e" % d["value"], "ms %.1f" % d["ms_per_step"], "n", d["config"]["particles"], "err", d["cpu_baseline"]["max_rel_err_gpu_vs_cpu"],
              "strong_ms", (d.get("strong") or {}).get("ms_per_step"))
    except Exception as e:
        print(f, "FAILED", e)
PY
