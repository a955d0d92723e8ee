"""ncu_lines.py <rep> <kernel regex> <mangled name substring> [n]: aggregates the ncu SASS page by CUDA source line,
using nvdisasm -g line info of the in-tree libba_cuda.so (the report's own CUDA view carries no metrics in csv mode)."""
import sys, csv, subprocess, io, re, os, collections
rep, kern, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
so = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "realsensecalibration_b200/csrc/libba_cuda.so")
import shutil
shutil.rmtree("/tmp/cub", ignore_errors=True)
os.makedirs("/tmp/cub", exist_ok=True)
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd="/tmp/cub", capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", "/tmp/cub/ba_cuda.sm_100a.cubin"], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and mangled in l)
lines, cur = [], ("?", 0)
for l in sass[start + 1:]:
    if l.startswith(".text.") or l.startswith("//-----"): 
        if lines: break
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*);", l)
    if m: lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
blk = out.split('"Kernel Name"')[1]
rows = list(csv.reader(io.StringIO("\n".join(blk.split("\n")[1:]))))
hdr = rows[0]
data = [r for r in rows[1:] if len(r) == len(hdr)]
assert len(data) == len(lines), (len(data), len(lines))
def num(r, n):
    try: return float(r[hdr.index(n)].replace(",", ""))
    except Exception: return 0.0
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.defaultdict(lambda: collections.Counter())
for r, (off, src, ins) in zip(data, lines):
    a = agg[src]
    a["samples"] += num(r, "# Samples"); a["wf"] += num(r, "L1 Wavefronts Shared"); a["ideal"] += num(r, "L1 Wavefronts Shared Ideal")
    a["inst"] += num(r, "Instructions Executed")
    for h in stalls: a[h[6:]] += num(r, h)
tot = sum(a["samples"] for a in agg.values())
print("samples", tot, "wavefronts", sum(a["wf"] for a in agg.values()), "ideal", sum(a["ideal"] for a in agg.values()), "inst", sum(a["inst"] for a in agg.values()))
srcfiles = {f: open(os.path.join(os.path.dirname(so), f)).read().split("\n") for f in os.listdir(os.path.dirname(so)) if f.endswith((".cuh", ".cu"))}
for src, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:topn]:
    st = sorted(((a[h[6:]], h[6:]) for h in stalls), reverse=True)[:3]
    text = srcfiles[src[0]][src[1] - 1].strip()[:90] if src[0] in srcfiles and 0 < src[1] <= len(srcfiles[src[0]]) else ""
    print("%5.1f%% wf %9d/%9d inst %8d %-34s %s:%d  %s" % (100 * a["samples"] / tot, a["wf"], a["ideal"], a["inst"],
          " ".join("%s:%d" % (n, v) for v, n in st if v > 0), src[0], src[1], text))
