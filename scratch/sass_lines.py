"""Static SASS attribution to source lines for one kernel of a cubin disassembled with `nvdisasm -g`.
usage: python scratch/sass_lines.py <nvdisasm -g output> <substring of the .text section name> [OPCODE ...]"""
import collections, re, sys
path, key, ops = sys.argv[1], sys.argv[2], set(sys.argv[3:])
cur = None; on = False
by_line = collections.Counter(); by_op = collections.Counter()
for l in open(path):
    if l.startswith(".text."):
        on = key in l
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        op = m.group(2).split(".")[0]
        by_op[op] += 1
        if not ops or op in ops: by_line[cur] += 1
print("static instructions:", sum(by_op.values()), dict(by_op.most_common(12)))
for k, v in by_line.most_common(40): print(k, v)
