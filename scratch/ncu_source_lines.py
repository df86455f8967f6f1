"""Per-source-line totals (instructions executed, stall samples) from `ncu --page source --csv --print-source cuda,sass`.
usage: python scratch/ncu_source_lines.py <csv> [min_share_percent]"""
import csv, sys, collections
path = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = list(csv.reader(open(path)))
cur_file = None; hdr = None
lines = []   # (file, line, src, inst, samples)
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "": continue            # SASS row
    try: ln = int(r[0])
    except ValueError: continue
    i_inst = hdr.index("Instructions Executed"); i_s = hdr.index("# Samples")
    lines.append((cur_file, ln, r[1].strip(), int(r[i_inst]) if r[i_inst].isdigit() else 0, int(r[i_s]) if r[i_s].isdigit() else 0))
ti = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print("total inst %d, samples %d" % (ti, ts))
for f, ln, src, inst, s in lines:
    if 100.0 * inst / ti >= thr or 100.0 * s / ts >= thr:
        print("%-18s %4d  inst %5.2f%%  samples %5.2f%%  %s" % (f, ln, 100.0 * inst / ti, 100.0 * s / ts, src[:100]))
