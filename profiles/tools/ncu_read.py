"""ncu_read.py <rep> [substr ...]: prints the raw-page metrics whose names contain any of the substrings, per kernel."""
import sys, csv, subprocess, io
rep = sys.argv[1]
subs = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct", "warps_active.avg.pct",
    "registers_per_thread", "pipe_fp64", "wavefronts_mem_shared.sum", "bank_conflicts_pipe_lsu_mem_shared.sum", "issue_active.avg.pct",
    "l1tex__throughput.avg.pct", "lts__t_sector_hit_rate", "smsp__inst_executed.sum", "shared_mem_per_block", "occupancy_limit", "lts__throughput.avg.pct",
    "issue_stalled", "achieved_occupancy", "waves_per_multiprocessor"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("==", r[hdr.index("Kernel Name")], r[hdr.index("Block Size")], r[hdr.index("Grid Size")])
    for i, h in enumerate(hdr):
        if any(s in h for s in subs):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if "issue_stalled" in h and ("pct" not in h and "ratio" not in h): continue
            if "issue_stalled" in h and v < 0.05: continue
            print("  %-110s %16s %s" % (h.split(".", 2)[-1] if h.count(".") > 2 else h, r[i], units[i]))
