"""Summarise an `ncu --page source --csv` export: instruction mix by opcode + stall reasons."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ci = {n: i for i, n in enumerate(h)}
ops = collections.Counter(); thr = collections.Counter(); samples = collections.Counter()
stall_cols = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
stalls = collections.Counter()
tot_inst = tot_thr = 0
for r in rows[2:]:
    if len(r) < len(h): continue
    src = r[ci['Source']].strip()
    toks = src.split()
    if not toks: continue
    op = toks[1] if toks[0].startswith('@') and len(toks) > 1 else toks[0]
    op = op.split('.')[0]
    try:
        n = int(r[ci['Instructions Executed']]); t = int(r[ci['Thread Instructions Executed']])
    except ValueError:
        continue
    ops[op] += n; thr[op] += t; tot_inst += n; tot_thr += t
    try: samples[op] += int(r[ci['# Samples']])
    except ValueError: pass
    for s in stall_cols:
        try: stalls[s] += int(r[ci[s]])
        except ValueError: pass
print("total warp-inst %d thread-inst %d avg threads %.2f" % (tot_inst, tot_thr, tot_thr / max(1, tot_inst)))
print("%-10s %12s %6s %8s %8s" % ("op", "warp-inst", "%", "avgthr", "samples%"))
ts = sum(samples.values()) or 1
for op, n in ops.most_common(28):
    print("%-10s %12d %6.2f %8.2f %8.2f" % (op, n, 100.0 * n / tot_inst, thr[op] / max(1, n), 100.0 * samples[op] / ts))
fp64 = sum(n for op, n in ops.items() if op in ('DFMA', 'DMUL', 'DADD', 'DSETP', 'DMNMX'))
print("fp64-pipe share of warp-inst: %.1f%%" % (100.0 * fp64 / tot_inst))
tt = sum(stalls.values()) or 1
print("stalls:", ", ".join("%s %.1f%%" % (k.replace('stall_', ''), 100.0 * v / tt) for k, v in stalls.most_common(8)))
