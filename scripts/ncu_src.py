"""Summarise an `ncu --page source --csv --print-source sass` dump: hottest instructions and their stall reasons."""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
recs = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: s = int(r[ix['# Samples']]); e = int(r[ix['Instructions Executed']])
    except ValueError: continue
    st = {h: int(r[ix[h]] or 0) for h in stalls}
    recs.append((s, e, r[ix['Address']], r[ix['Source']], st))
tot = sum(r[0] for r in recs)
agg = collections.Counter()
for s, e, a, src, st in recs:
    for h, v in st.items(): agg[h] += v
print("total samples", tot, {h.replace('stall_', ''): v for h, v in agg.most_common(8)})
for i, (s, e, a, src, st) in enumerate(recs):
    pass
order = sorted(range(len(recs)), key=lambda i: -recs[i][0])[:top]
for i in sorted(order):
    s, e, a, src, st = recs[i]
    best = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{i:5d} {100*s/tot:5.1f}% ex={e:9d} {src[:70]:70s} {[(k.replace('stall_',''),v) for k,v in best]}")
