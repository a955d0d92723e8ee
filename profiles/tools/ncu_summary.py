"""ncu_summary.py <rep> : first captured launch of every kernel in an `ncu --set full` report, reduced to the metrics
DESIGN.md quotes.  Also prints the DRAM bytes per launch as JSON (for profiles/traffic.json)."""
import sys, csv, subprocess, io, json
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
want = [("time", "gpu__time_duration.sum"), ("dram_read", "dram__bytes_read.sum"), ("dram_write", "dram__bytes_write.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("grid", "launch__grid_size"), ("block", "launch__block_size"),
        ("regs", "launch__registers_per_thread"), ("dyn_smem", "launch__shared_mem_per_block_dynamic"),
        ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("issue_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        ("fp64_pipe_pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"), ("l1tex_pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
        ("smem_wavefronts", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"), ("smem_conflicts", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        ("l2_hit_pct", "lts__t_sector_hit_rate.pct"), ("warp_insts", "smsp__inst_executed.sum"),
        ("stall_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
        ("stall_short_scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
        ("stall_barrier", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio")]
def find(name, row):
    hit = -1
    for i, h in enumerate(hdr):
        if h == name or h.endswith("." + name) or h.split(".", 2)[-1] == name:
            if row[i].strip(): return i
            hit = i
    return hit
seen, traffic = set(), {}
print("# ncu --set full --clock-control none, first captured launch of each kernel; source: %s" % rep.split("/")[-1])
for r in rows[2:]:
    k = r[hdr.index("Kernel Name")]
    if k in seen: continue
    seen.add(k)
    print("\n" + k)
    vals = {}
    for label, name in want:
        i = find(name, r)
        if i < 0: continue
        print("  %-22s %s %s" % (label, r[i], units[i]))
        vals[label] = (r[i], units[i])
    def nbytes(v):
        x, u = v; x = float(x.replace(",", ""))
        return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if "dram_read" in vals and "dram_write" in vals: traffic[k] = int(nbytes(vals["dram_read"]) + nbytes(vals["dram_write"]))
print("\n# dram bytes per launch: " + json.dumps(traffic))
