"""Summarise `ncu -i REPORT --page source --csv` (stdin) into per-opcode stall shares."""
import collections
import csv
import sys

rows = list(csv.reader(sys.stdin))
kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
si = hdr.index("Source")
ai = hdr.index("Warp Stall Sampling (All Samples)")
ei = hdr.index("Instructions Executed")
data = [r for r in data if r[ai].isdigit()]
seen = set()
data = [r for r in data if not (r[0] in seen or seen.add(r[0]))]
tot = sum(int(r[ai]) for r in data)
byop = collections.Counter()
execop = collections.Counter()
for r in data:
    tok = r[si].strip().split()
    if not tok:
        continue
    op = tok[1] if tok[0].startswith("@") and len(tok) > 1 else tok[0]
    op = op.split(".")[0]
    byop[op] += int(r[ai])
    execop[op] += int(r[ei]) if r[ei].isdigit() else 0
texec = sum(execop.values())
print(kernel)
print("total stall samples", tot, " warp-instructions executed", texec)
print("%-10s %10s %7s %16s %7s" % ("opcode", "samples", "share", "warp-instr", "share"))
for op, c in byop.most_common(16):
    print("%-10s %10d %6.1f%% %16d %6.1f%%" % (op, c, 100 * c / tot, execop[op], 100 * execop[op] / texec))
print()
print("top 12 instructions by stall samples:")
for r in sorted(data, key=lambda r: -int(r[ai]))[:12]:
    print("%8s  %s" % (r[ai], r[si].strip()))
