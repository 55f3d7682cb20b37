"""Aggregate an ncu SASS-level source page by CUDA source line.
  python tools/ncu_by_line.py REPORT.ncu-rep CUBIN KERNEL_MANGLED [launch index]
Uses `ncu --page source --csv --print-source sass` for the per-instruction counters and
`nvdisasm -g` of the cubin for the instruction -> file:line map (matched by instruction order)."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict

rep, cubin, kern = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(txt)):
    if row and row[0] == "Kernel Name":
        cur = dict(name=row[1], rows=[])
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(row)
blk = blocks[which]
hdr = blk["rows"][0]
rows = [r for r in blk["rows"][1:] if len(r) == len(hdr)]
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kern + ":"))
lines, line = [], None
for l in dis[start + 1:]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        line = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append((line, l.strip()))
    if l.startswith("\t.section") or l.startswith(".text.") and kern not in l:
        break
n = min(len(lines), len(rows))
ie, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
te = hdr.index("Thread Instructions Executed")
agg = defaultdict(lambda: [0, 0, 0])
for (ln, _), r in zip(lines[:n], rows[:n]):
    a = agg[ln]
    a[0] += int(r[ie]); a[1] += int(r[sm]); a[2] += int(r[te])
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print("kernel", blk["name"][:60], "sass", len(rows), "disasm", len(lines), "warp-instr", tot, "samples", tots)
src = {}
key = (lambda kv: -kv[1][0]) if len(sys.argv) > 5 and sys.argv[5] == "instr" else (lambda kv: -kv[1][1])
for ln, a in sorted(agg.items(), key=key)[:40]:
    if ln and ln[0] not in src:
        try:
            src[ln[0]] = open("pyellipticfd_b200/csrc/" + ln[0]).read().splitlines()
        except OSError:
            src[ln[0]] = []
    text = src[ln[0]][ln[1] - 1].strip()[:90] if ln and len(src[ln[0]]) >= ln[1] else ""
    print("%6.2f%% instr %6.2f%% samples  thr/instr %4.1f  %s:%s  %s" % (
        100 * a[0] / tot, 100 * a[1] / max(tots, 1), a[2] / max(a[0], 1), ln[0] if ln else "?", ln[1] if ln else "?", text))
