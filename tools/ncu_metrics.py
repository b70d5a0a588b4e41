"""Print the headline metrics of every launch in an .ncu-rep (read here, no GPU needed).
Usage: python tools/ncu_metrics.py gpurun_out/prof_x.ncu-rep [extra_metric_substring ...]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "sm__cycles_active.avg",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__waves_per_multiprocessor", "gpc__cycles_elapsed.max", "sm__cycles_active.max", "launch__sm_count"]
for i, h in enumerate(hdr):
    if h in want or any(e in h for e in extra) or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        vals = [r[i] for r in data]
        if "issue_stalled" in h:
            if max(float(v.replace(",", "") or 0) for v in vals) < 0.3:
                continue
            h = h.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "")
        print(f"{h[:70]:70s} {units[i]:12s} {vals}")
