"""Instruction mix / pipe utilisation of one ncu --set full capture: python scratch/ncu_mix.py <file.ncu-rep> <n_events>"""
import collections, csv, io, subprocess, sys
rep, n_ev = sys.argv[1], float(sys.argv[2])
def ncu(page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(ncu("source"))))
sh = rows[1]; ci = {n: i for i, n in enumerate(sh)}
ops = collections.Counter(); tot = 0; stalls = collections.Counter()
stall_cols = [n for n in sh if n.startswith("stall_") and "Not Issued" not in n]
for r in rows[2:]:
    if len(r) < len(sh): continue
    toks = r[ci["Source"]].split()
    if not toks: continue
    op = (toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0])
    op = ".".join(op.split(".")[:2]) if op.split(".")[0] in ("LDS", "STS", "LDG", "STG", "MUFU", "F2F", "F2I", "I2F", "SHFL") else op.split(".")[0]
    try: n = int(r[ci["Instructions Executed"]])
    except ValueError: continue
    ops[op] += n; tot += n
    for s in stall_cols:
        try: stalls[s] += int(r[ci[s]])
        except ValueError: pass
we = n_ev / 32
print("warp-instructions per warp-event: %.1f" % (tot / we))
for op, n in ops.most_common(40): print("%-14s %8.1f %5.1f%%" % (op, n / we, 100 * n / tot))
ts = sum(stalls.values()) or 1
print("stalls:", ", ".join("%s %.1f%%" % (k.replace("stall_", ""), 100.0 * v / ts) for k, v in stalls.most_common(10)))
raw = list(csv.reader(io.StringIO(ncu("raw"))))
h, u, r = raw[0], raw[1], raw[2]
for k in ("Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
          "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
          "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum"):
    if k in h: print(k, r[h.index(k)][:90], u[h.index(k)])
