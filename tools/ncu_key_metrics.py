"""Key metrics of `ncu --set full` reports (read with `ncu -i REP --page raw --csv`) -> text table for profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__cluster_size",
        "launch__occupancy_limit_shared_mem", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== %s   [%s]" % (r[hdr.index("Kernel Name")][:110], rep.split("/")[-1]))
        for k in KEYS:
            if k in hdr:
                print("   %-75s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        stalls = sorted(((float(r[i] or 0), h) for i, h in enumerate(hdr)
                         if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")
                         and "not_issued" not in h), reverse=True)[:6]
        print("   top warp stalls per issue-active cycle: " +
              ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
                        for v, h in stalls))
        print()
