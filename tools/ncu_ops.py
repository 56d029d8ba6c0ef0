"""Executed warp instructions per SASS opcode from `ncu --page source --csv --print-source sass`."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for i, r in enumerate(rows):
    if 'Instructions Executed' in r:
        hdr, start = r, i + 1
        break
ii, si, ti = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('Thread Instructions Executed')
ops, tot, tt = defaultdict(int), 0, 0
for r in rows[start:]:
    if len(r) <= ii:
        continue
    try:
        n, t_ = int(r[ii] or 0), int(r[ti] or 0)
    except ValueError:
        continue
    t = r[si].split()
    if not t:
        continue
    op = t[1] if t[0].startswith('@') else t[0]
    ops[op.split('.')[0]] += n
    tot += n
    tt += t_
print(f"total warp instructions {tot:.4e}, threads per instruction {tt / max(tot, 1):.1f}")
for k, v in sorted(ops.items(), key=lambda x: -x[1])[:top]:
    print(f"{k:12s} {v:12d} {100 * v / tot:5.1f}%")
