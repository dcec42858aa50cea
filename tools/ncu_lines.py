"""Attribute ncu per-instruction counters to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-name-substring> [top]

ncu's `--page source --csv` lists SASS instructions in program order with executed counts and stall samples;
`nvdisasm -g` lists the same instructions with `//## File ..., line N` markers.  Zipping the two gives a per-line
profile without a GUI."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
data = rows[2:]
iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
# find the kernel section
start = None
for k, line in enumerate(sass):
    if line.startswith(".text.") and pat in line:
        start = k
        break
assert start is not None, "kernel not found in cubin"
lines = []
cur = None
instr_re = re.compile(r"^\s+(/\*[0-9a-f]+\*/)?\s+(@!?U?P\d+\s+)?[A-Z][A-Z0-9_.]+")
for line in sass[start + 1:]:
    if line.startswith(".text.") or line.startswith("//----"):
        if lines:
            break
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"^\s*/\*[0-9a-f]{4,}\*/", line):
        lines.append(cur)
print(f"ncu instructions: {len(data)}, nvdisasm instructions: {len(lines)}")
n = min(len(data), len(lines))
agg_e, agg_s = defaultdict(int), defaultdict(int)
for k in range(n):
    agg_e[lines[k]] += int(data[k][iE])
    agg_s[lines[k]] += int(data[k][iSm])
te, ts = sum(agg_e.values()), sum(agg_s.values())
src_cache = {}
def src(f, l):
    import os
    path = os.path.join("/root/repo/carmpc_b200/csrc", f)
    if path not in src_cache:
        try:
            src_cache[path] = open(path).read().split("\n")
        except OSError:
            src_cache[path] = []
    t = src_cache[path]
    return t[l - 1].strip()[:90] if 0 < l <= len(t) else ""
print(f"{'file:line':28s} {'exec%':>6s} {'stall%':>6s}  source")
for key, e in sorted(agg_e.items(), key=lambda kv: -agg_s[kv[0]])[:top]:
    if key is None:
        continue
    print(f"{key[0] + ':' + str(key[1]):28s} {e / te * 100:6.2f} {agg_s[key] / ts * 100:6.2f}  {src(*key)}")
