#!/usr/bin/env python3
"""Attribute an ncu source-page export (SASS view) to CUDA source lines.
    ncu -i rep.ncu-rep --page source --csv > sass.csv
    nvdisasm -g -c kernels.cubin > kernels.dis
    python tools/ncu_lines.py sass.csv kernels.dis <mangled kernel name> [top]
The i-th instruction of the ncu listing is the i-th instruction of the function in the disassembly; nvdisasm's
`//## File "...", line N` markers give the source line of what follows (innermost inline frame)."""
import csv
import re
import sys

sass_csv, dis, fn = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = []
cur = None
infn = False
for ln in open(dis):
    if ln.startswith(".text."):
        infn = ln.strip() == f".text.{fn}:"
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(open(sass_csv)))
hdr = None
data = []
for r in rows:
    if r and r[0] == "Kernel Name":
        if data:
            break
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        data.append(r)
assert len(data) == len(lines), (len(data), len(lines))
iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
agg = {}
for r, l in zip(data, lines):
    a = agg.setdefault(l, [0, 0])
    a[0] += int(r[iex] or 0)
    a[1] += int(r[ismp] or 0)
tex = sum(a[0] for a in agg.values())
tsm = sum(a[1] for a in agg.values())
print(f"total instructions {tex}, samples {tsm}")
src = {}
for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        try:
            src[f] = open(f"sculptmate_b200/csrc/{f}").read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][n - 1].strip()[:90] if n - 1 < len(src[f]) else ""
    print(f"{f}:{n:5d}  inst {100 * a[0] / tex:5.1f}%  samples {100 * a[1] / max(tsm, 1):5.1f}%  {text}")

# optional: aggregate by ranges of the kernel body ("name:first_line" ...), helper-function lines inherit the range of the
# last body line seen before them in SASS order
if len(sys.argv) > 5:
    marks = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[5:]]
    body_lo = min(m[1] for m in marks)
    cur_rng = None
    rng = {}
    for r, l in zip(data, lines):
        if l and l[0] == "mcubes.cu" and l[1] >= body_lo:
            cur_rng = [m[0] for m in marks if m[1] <= l[1]][-1]
        a = rng.setdefault(cur_rng, [0, 0])
        a[0] += int(r[iex] or 0)
        a[1] += int(r[ismp] or 0)
    for k, a in rng.items():
        print(f"range {k}: inst {100 * a[0] / tex:5.1f}%  samples {100 * a[1] / max(tsm, 1):5.1f}%")
