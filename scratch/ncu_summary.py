"""profiles/<name>_ncu_summary.md + profiles/ncu_traffic.json entry from one `ncu --set full` capture of the fused kernel.
usage: python scratch/ncu_summary.py <file.ncu-rep> <n_events> <name> <traffic key> "<one-line description>" """
import collections, csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, n_ev, name, key, desc = sys.argv[1], float(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(ncu("raw"))))
h, u, r = raw[0], raw[1], raw[2]
val = lambda k: r[h.index(k)] if k in h else "n/a"
unit = lambda k: u[h.index(k)] if k in h else ""
num = lambda k: float(val(k).replace(",", ""))
to_bytes = lambda k: num(k) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit(k)]
dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
rows = list(csv.reader(io.StringIO(ncu("source"))))
sh = rows[1]; ci = {n: i for i, n in enumerate(sh)}
ops = collections.Counter(); thr = collections.Counter(); stalls = collections.Counter(); tot = 0
stall_cols = [n for n in sh if n.startswith("stall_") and "Not Issued" not in n]
for row in rows[2:]:
    if len(row) < len(sh): continue
    toks = row[ci["Source"]].split()
    if not toks: continue
    op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]).split(".")[0]
    try: n = int(row[ci["Instructions Executed"]]); t = int(row[ci["Thread Instructions Executed"]])
    except ValueError: continue
    ops[op] += n; thr[op] += t; tot += n
    for s in stall_cols:
        try: stalls[s] += int(row[ci[s]])
        except ValueError: pass
we = n_ev / 32
flop64 = (2 * thr["DFMA"] + thr["DMUL"] + thr["DADD"]) / n_ev
flop32 = (2 * thr["FFMA"] + thr["FMUL"] + thr["FADD"] + 4 * thr["FFMA2"] + 2 * thr["FMUL2"] + 2 * thr["FADD2"]) / n_ev
L = []
L.append("# %s: `%s`\n" % (name, val("Kernel Name").split("(")[0]))
L.append(desc + "\n")
L.append("| metric | value |\n|---|---|")
for k in ("gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
          "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
          "smsp__thread_inst_executed_per_inst_executed.ratio"):
    L.append("| %s | %s %s |" % (k, val(k), unit(k)))
L.append("| DRAM bytes / event | %.1f |" % (dram / n_ev))
L.append("| warp instructions / warp-event | %.0f |" % (tot / we))
L.append("| executed FP64 FLOP / event (2 DFMA + DMUL + DADD) | %.0f |" % flop64)
L.append("| executed FP32 FLOP / event (2 FFMA + FMUL + FADD, packed x2) | %.0f |" % flop32)
L.append("\nSASS mix (warp instructions per warp-event):\n\n```")
for op, n in ops.most_common(26): L.append("%-10s %8.1f %6.2f%%" % (op, n / we, 100.0 * n / tot))
ts = sum(stalls.values()) or 1
L.append("stalls: " + ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / ts) for k, v in stalls.most_common(8)))
L.append("```")
open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % name), "w").write("\n".join(L) + "\n")
tj = os.path.join(ROOT, "profiles", "ncu_traffic.json")
d = json.load(open(tj)) if os.path.exists(tj) else {}
d[key] = {"dram_bytes_per_event": dram / n_ev, "fp64_pipe_active_pct": num("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
          "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
          "executed_fp64_flop_per_event": flop64, "executed_fp32_flop_per_event": flop32,
          "warp_instructions_per_warp_event": tot / we, "events_in_capture": n_ev,
          "source": "profiles/%s_ncu_summary.md" % name}
json.dump(d, open(tj, "w"), indent=1)
print("\n".join(L[:30]))
