"""Top stall sites of a kernel from `ncu --page source --csv --print-source sass` output:  python tools/ncu_hot.py rep.ncu-rep [N]"""
import csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] in ('Address', 'Line No', '#'))
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
samp = col['# Samples']; src = col['Source']; ex = col['Instructions Executed']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for r in rows[hi + 1:]:
    if len(r) <= samp: continue
    try: s = int(r[samp])
    except ValueError: continue
    data.append((s, r))
tot = sum(s for s, _ in data)
print('total samples', tot, 'instructions', len(data))
agg = {h: 0 for h in stalls}
for s, r in data:
    for h in stalls:
        try: agg[h] += int(r[col[h]])
        except ValueError: pass
print('stall totals:', {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for idx, (s, r) in enumerate(data):
    pass
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:n]
for i in sorted(order):
    s, r = data[i]
    top = sorted(((int(r[col[h]] or 0), h) for h in stalls), reverse=True)[:2]
    print(f'{i:5d} {s:7d} {100*s/tot:5.1f}%  ex={r[ex]:>9}  {r[src][:70]:70s} {top}')
