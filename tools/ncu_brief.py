"""Print the handful of ncu metrics used in DESIGN.md / profiles/ from a .ncu-rep (first kernel): python tools/ncu_brief.py file.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units = r[0], r[1]
for row in r[2:]:
    def g(k):
        i = hdr.index(k); return row[i] + " " + units[i]
    for k in ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
              "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "launch__waves_per_multiprocessor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
              "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed"]:
        try: print(f"  {k} = {g(k)}")
        except ValueError: pass
    st = []
    for i, h in enumerate(hdr):
        if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
            try: st.append((float(row[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
            except ValueError: pass
    print("  stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))
