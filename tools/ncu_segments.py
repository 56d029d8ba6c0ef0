"""Segment-level view of an `ncu --page source --csv` dump: runs of SASS instructions with the same
execution count, their share of instructions / stall samples and the top stall reasons."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
si = hdr.index('Source'); ii = hdr.index('Instructions Executed')
segs = []
for k, r in enumerate(rows[2:]):
    try: ie = int(r[ii] or 0)
    except Exception: continue
    st = collections.Counter({hdr[i].replace('stall_', ''): int(r[i] or 0) for i in cols})
    if segs and abs(segs[-1]['ie'] - ie) <= 0.08 * max(ie, 1) + 100:
        s = segs[-1]; s['n'] += 1; s['inst'] += ie; s['st'] += st; s['mufu'] += ('MUFU' in r[si])
    else:
        segs.append({'k': k, 'ie': ie, 'n': 1, 'inst': ie, 'st': st, 'mufu': int('MUFU' in r[si]), 'src': r[si][:40]})
ti = sum(s['inst'] for s in segs); ts = sum(sum(s['st'].values()) for s in segs)
tot = collections.Counter()
for s in segs: tot += s['st']
print(f"warp instructions {ti:.3e}, samples {ts}; stall mix " + ", ".join(f"{k} {100*v/ts:.1f}%" for k, v in tot.most_common(7)))
for s in segs:
    smp = sum(s['st'].values())
    if s['inst'] > 0.004 * ti or smp > 0.004 * ts:
        print(f"{s['k']:5d} n={s['n']:4d} exec={s['ie']:9d} inst={100*s['inst']/ti:5.1f}% samp={100*smp/ts:5.1f}% mufu={s['mufu']:3d} {dict(s['st'].most_common(3))} {s['src']}")
