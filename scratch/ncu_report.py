"""Turn a gpurun cycle (scratch/gpu_cycle.sh <tag>) into committed evidence under profiles/:
   profiles/<round>_<tag>_ncu_summary.md, _launches.csv, _bench.json and profiles/ncu_traffic.json."""
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]                     # e.g. r1_j
name = sys.argv[2] if len(sys.argv) > 2 else tag
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

def ncu(page):
    return subprocess.run(["ncu", "-i", os.path.join(G, "prof_%s.ncu-rep" % tag), "--page", page, "--csv"],
                          capture_output=True, text=True).stdout

raw = list(csv.reader(io.StringIO(ncu("raw"))))
h, u, r = raw[0], raw[1], raw[2]
val = lambda k: r[h.index(k)] if k in h else "n/a"
unit = lambda k: u[h.index(k)] if k in h else ""
num = lambda k: float(val(k).replace(",", ""))
def to_bytes(k):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit(k)]
    return num(k) * f
n_events = int(float(sys.argv[3])) if len(sys.argv) > 3 else 12_000_000
kname = val("Kernel Name").split("(")[0]
dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")

src = list(csv.reader(io.StringIO(ncu("source"))))
sh = src[1]; ci = {n: i for i, n in enumerate(sh)}
ops = collections.Counter(); thr = collections.Counter(); stalls = collections.Counter()
stall_cols = [n for n in sh if n.startswith("stall_") and "Not Issued" not in n]
tot = 0
for row in src[2:]:
    if len(row) < len(sh):
        continue
    toks = row[ci["Source"]].split()
    if not toks:
        continue
    op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
    try:
        n = int(row[ci["Instructions Executed"]]); t = int(row[ci["Thread Instructions Executed"]])
    except ValueError:
        continue
    ops[op] += n; thr[op] += t; tot += n
    for s in stall_cols:
        try: stalls[s] += int(row[ci[s]])
        except ValueError: pass
fp64_thread = {k: thr[k] for k in ("DFMA", "DMUL", "DADD")}
flop_event = (2 * fp64_thread["DFMA"] + fp64_thread["DMUL"] + fp64_thread["DADD"]) / n_events
inst_event = sum(thr.values()) / n_events
ts = sum(stalls.values()) or 1

lines = []
L = lines.append
L("# %s: `%s` (ncu --set full --clock-control none, B200)\n" % (name, kname))
L("One launch = one template over %d events (12 flavour containers), PREM_12layer, nufit 2.0 NH; command in" % n_events)
L("`scratch/gpu_cycle.sh` (bench.py --events-per-gpu 1.2e7 --steps 1 --warmup 3 under ncu, NVTX range `timed`).\n")
L("| metric | value |\n|---|---|")
for k in ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
          "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
          "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"):
        L("| %s | %s %s |" % (k, val(k), unit(k)))
L("| DRAM bytes / event | %.1f (algorithmic 44) |" % (dram / n_events))
L("| executed thread instructions / event | %.0f |" % inst_event)
L("| executed FP64 FLOP / event (2 DFMA + DMUL + DADD) | %.0f (reference arithmetic: ~13 585) |" % flop_event)
L("\nSASS mix (warp instructions):\n\n```")
L("%-10s %12s %6s %8s" % ("op", "warp-inst", "%", "avg thr"))
for op, n in ops.most_common(24):
    L("%-10s %12d %6.2f %8.2f" % (op, n, 100.0 * n / tot, thr[op] / max(1, n)))
L("stalls: " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / ts) for k, v in stalls.most_common(8)))
L("```")
open(os.path.join(P, "%s_ncu_summary.md" % name), "w").write("\n".join(lines) + "\n")

# launch list of the timed region
rows = list(csv.reader(l for l in open(os.path.join(G, "launches_%s.csv" % tag)) if l.startswith('"')))
lh = rows[0]; ki, vi, ui = lh.index("Kernel Name"), lh.index("Metric Value"), lh.index("Metric Unit")
t = collections.Counter(); c = collections.Counter()
for row in rows[1:]:
    try: v = float(row[vi].replace(",", ""))
    except ValueError: continue
    v = v / 1e3 if row[ui] == "ns" else v * 1e3 if row[ui] == "ms" else v
    k = row[ki].split("(")[0]; t[k] += v; c[k] += 1
with open(os.path.join(P, "%s_launches.csv" % name), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --nvtx-include timed/ : kernels of the timed region (2 steps, 1.2e7 events)\n")
    f.write("kernel,launches,total_us,share_pct,avg_us\n")
    for k, v in t.most_common():
        f.write("%s,%d,%.1f,%.1f,%.1f\n" % (k, c[k], v, 100 * v / sum(t.values()), v / c[k]))
for src_n, dst_n in (("bench_%s.json" % tag, "%s_bench.json" % name), ("bench_ref_%s.json" % tag, "%s_bench_ref.json" % name)):
    src_f = os.path.join(G, src_n)
    if os.path.exists(src_f):
        open(os.path.join(P, dst_n), "w").write(open(src_f).read())
tj = os.path.join(P, "ncu_traffic.json")
d = json.load(open(tj)) if os.path.exists(tj) else {}
d["reweight_hist_kernel<double>"] = {
    "dram_bytes_per_event": dram / n_events, "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
    "executed_fp64_flop_per_event": flop_event, "source": "profiles/%s_ncu_summary.md" % name}
json.dump(d, open(tj, "w"), indent=1)
print(open(os.path.join(P, "%s_ncu_summary.md" % name)).read())
print(open(os.path.join(P, "%s_launches.csv" % name)).read())
