"""Text summary of an ncu --set full report for profiles/: per kernel the launch shape, duration, DRAM bytes, pipe
utilisation, issue statistics and the top stall reasons.
usage: python tools/ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_global_red.sum"]
stall = [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]
print("# " + " ".join(sys.argv[2:]))
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("\n== " + d["Kernel Name"][:90])
    for k in keys:
        if k in d and d[k] != "":
            print(f"{k:<75} {d[k]:>16} {units[hdr.index(k)]}")
    st = sorted(((float(d[h] or 0), h.split("issue_stalled_")[1].split("_per_")[0]) for h in stall), reverse=True)[:8]
    print("stall reasons (warps per issue): " + ", ".join(f"{n}={v:.2f}" for v, n in st))
