"""Print key metrics + stall breakdown of an .ncu-rep (run here, no GPU)."""
import csv, subprocess, sys
for f in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    for v in rows[2:]:
        d = {h[i]: (v[i], u[i]) for i in range(len(h))}
        print("==", f, d.get("Kernel Name", ("", ""))[0][:60])
        keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
                "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
                "launch__grid_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
                "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
                "sm__cycles_elapsed.max"]
        for k in keys:
            if k in d:
                print("  %-62s %s %s" % (k, d[k][0], d[k][1]))
        st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(d[k][0]))
              for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
              and "not_issued" not in k]
        print("  stalls:", ", ".join("%s %.2f" % kv for kv in sorted(st, key=lambda x: -x[1])[:9]))
