"""Per CUDA source line totals from `ncu --page source --csv --print-source cuda,sass`:
warp instructions executed and stall samples, top N lines per file."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur, files = None, []
for r in rows:
    if r and r[0] == 'File Path':
        cur = {'name': r[1], 'hdr': None, 'rows': []}; files.append(cur); continue
    if cur is None or not r: continue
    if r[0] == 'Line No': cur['hdr'] = r; continue
    if r[0] == 'Function Name': continue
    cur['rows'].append(r)
grand = 0
out = []
for f in files:
    h = f['hdr']
    if not h: continue
    ii, si = h.index('Instructions Executed'), h.index('# Samples')
    for r in f['rows']:
        if len(r) <= ii or not r[0].isdigit(): continue   # source-line rows carry the line number
        try: ie, sm = int(r[ii] or 0), int(r[si] or 0)
        except ValueError: continue
        grand += ie
        out.append((ie, sm, f['name'].split('/')[-1], r[0], r[1].strip()[:100]))
print(f"total warp instructions {grand:.3e}")
for ie, sm, fn, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100*ie/grand:5.1f}% {sm:5d}  {fn}:{ln:>4s}  {src}")
