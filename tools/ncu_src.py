"""Summarise an `ncu --page source --csv` dump: instructions executed and stall samples per
SASS segment (segments split at BAR.SYNC / user-given addresses)."""
import csv, sys
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
ai, si, ni, ii, ti = (hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed"))
data = [(r[ai], r[si], int(r[ni] or 0), int(r[ii] or 0), int(r[ti] or 0)) for r in rows[2:] if len(r) > ii]
tot_i = sum(d[3] for d in data); tot_s = sum(d[2] for d in data)
print(f"total warp-inst {tot_i:.3e}  samples {tot_s}")
seg_start = 0
def flush(a, b, label):
    i = sum(d[3] for d in data[a:b]); s = sum(d[2] for d in data[a:b]); t = sum(d[4] for d in data[a:b])
    mufu = sum(d[3] for d in data[a:b] if "MUFU" in d[1])
    print(f"  [{a:4d},{b:4d}) {label:28s} inst {i:.3e} ({100*i/tot_i:5.1f}%)  samples {100*s/max(tot_s,1):5.1f}%  lanes/inst {t/max(i,1):5.1f}  mufu {mufu:.2e}")
for k, d in enumerate(data):
    if "BAR.SYNC" in d[1] or "EXIT" in d[1]:
        flush(seg_start, k + 1, f"..{d[0][-5:]} {d[1].split()[0]}")
        seg_start = k + 1
if seg_start < len(data): flush(seg_start, len(data), "tail")
if len(sys.argv) > 2:
    top = sorted(data, key=lambda d: -d[2])[: int(sys.argv[2])]
    for d in top: print(f"   {d[0][-5:]} samples {d[2]:6d} inst {d[3]:.2e}  {d[1][:70]}")
