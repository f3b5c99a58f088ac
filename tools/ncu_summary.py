"""Summarise an .ncu-rep (raw page) into a small CSV of the metrics DESIGN.md cites.
   python tools/ncu_summary.py in.ncu-rep out.csv "comment line"  """
import csv, io, subprocess, sys
KEEP = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "sm__warps_active.avg.per_cycle_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fp32_pred_on.sum",
        "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_fmaheavy.sum", "smsp__inst_executed_pipe_fmalite.sum",
        "smsp__inst_executed_pipe_xu.sum", "smsp__inst_executed_pipe_lsu.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum"]
STALL = "smsp__average_warps_issue_stalled_"
rep, out, comment = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
with open(out, "w") as f:
    f.write("# " + comment + "\n")
    f.write("metric,value,unit\n")
    f.write(f"Kernel Name,\"{d['Kernel Name'][0]}\",\n")
    for k in KEEP:
        if k in d:
            f.write(f"{k},{d[k][0]},{d[k][1]}\n")
    def num(k):
        try:
            return float(d[k][0].replace(",", ""))
        except Exception:
            return None
    # hardware-counted FP32 work.  The inner loops are packed FFMA2/FADD2/FMUL2, which the per-opcode thread counters
    # (smsp__sass_thread_inst_executed_op_{ffma,fadd,fmul}_pred_on) do NOT see -- they read ~0 for these kernels -- so the
    # count comes from the pipe: one packed warp instruction on the FMA pipe = 2 lanes x 32 threads = 64 lane-ops, and the
    # pipe retires 128 lane-ops per clock per SM.  (Kernels with scalar FP32 code also list the per-opcode counters.)
    fma, cyc, sms = num("smsp__inst_executed_pipe_fma.sum"), num("sm__cycles_elapsed.avg"), 148
    scal = [num("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % o) or 0.0 for o in ("fadd", "fmul", "ffma")]
    if fma and cyc:
        packed_lane = fma * 64.0 - sum(scal)      # scalar FMA-pipe instructions carry 32 lanes, not 64
        lane = packed_lane if sum(scal) < 0.01 * fma * 32 else fma * 32.0 + (fma * 32.0 - sum(scal))
        f.write(f"derived: FMA-pipe warp instructions x 64 lanes = FP32 lane-ops (packed FFMA2/FADD2/FMUL2),{fma * 64.0:.0f},\n")
        if sum(scal) < 0.01 * fma * 32:
            f.write(f"derived: hardware FP32 fraction = lane-ops / (128 lanes x {sms} SMs x elapsed cycles),{fma * 64.0 / (128.0 * sms * cyc):.4f},\n")
        xu = num("smsp__inst_executed_pipe_xu.sum")
        if xu:
            f.write(f"derived: XU fraction = MUFU lane-ops / (16 per clock x {sms} SMs x elapsed cycles),{xu * 32 / (16.0 * sms * cyc):.4f},\n")
    for k in sorted(d):
        if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and float(d[k][0] or 0) >= 0.1:
            f.write(f"{k},{d[k][0]},{d[k][1]}\n")
print(open(out).read())
