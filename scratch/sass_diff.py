"""Which kernels changed between two builds?  Compares the SASS instruction streams (addresses and encodings stripped)
of every function in two libraries / cuobjdump dumps.  Used in the build container (no GPU needed) to check that an edit
meant for one kernel left the tuned ones byte-for-byte alone before any GPU time is spent.

usage: python scratch/sass_diff.py <old.so | old_sass.txt> <new.so | new_sass.txt>"""
import re, subprocess, sys


def load(path):
    text = open(path, errors="ignore").read() if path.endswith(".txt") else \
        subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out, cur = {}, None
    for line in text.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur is not None:
            mm = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
            if mm:
                out[cur].append(mm.group(1))
    return out


if __name__ == "__main__":
    a, b = load(sys.argv[1]), load(sys.argv[2])
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:140]
    for k in sorted(a):
        if k not in b:
            print("REMOVED ", demangle(k))
        elif a[k] != b[k]:
            print("CHANGED  %5d -> %5d instructions  %s" % (len(a[k]), len(b[k]), demangle(k)))
    for k in sorted(b):
        if k not in a:
            print("NEW      %5d instructions  %s" % (len(b[k]), demangle(k)))
    print("%d functions before, %d after" % (len(a), len(b)))
