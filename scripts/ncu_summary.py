"""Extract the judged subset of an `ncu --set full` report of the MH kernel into a text summary (profiles/<tag>_mh_kernel_ncu_full.txt).
    python scripts/ncu_summary.py gpurun_out/r01_v7_prof_mh.ncu-rep > profiles/r01_v7_mh_kernel_ncu_full.txt"""
import csv
import subprocess
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "smsp__thread_inst_executed_per_inst_executed", "sm__warps_active.avg.pct", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem", "sm__cycles_elapsed.avg", "sm__inst_executed.sum.p",
        "sm__inst_executed_pipe_fp64", "sm__inst_executed_pipe_xu.avg", "sm__pipe_fp64_cycles_active.avg", "sm__issue_active.avg",
        "sm__warps_active.avg.per_cycle_active", "smsp__average_warps_issue_stalled", "smsp__issue_active.avg", "smsp__warps_eligible.avg",
        "smsp__inst_executed.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct", "gpu__compute_memory_throughput.avg.pct")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    if (sys.argv[2] if len(sys.argv) > 2 else "mh_kernel") not in d.get("Kernel Name", ""):
        continue
    print("%-103s %s" % ("Kernel Name", d["Kernel Name"]))
    for k, u in sorted(zip(hdr, units)):
        if any(k.startswith(p) for p in KEEP) and d.get(k, "") not in ("", "0", "0.000000"):
            print("%-90s %-12s %s" % (k, u, d[k]))
