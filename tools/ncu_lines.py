"""Per-CUDA-source-line stall samples of a kernel:  python tools/ncu_lines.py rep.ncu-rep [file-substring] [N]
(needs a capture made with --import-source on and a build with -lineinfo)"""
import csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
want = sys.argv[2] if len(sys.argv) > 2 else ''
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur_file, hdr, agg, tot = None, None, {}, 0
for r in csv.reader(raw.splitlines()):
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1]; continue
    if r[0] == 'Line No':
        hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or r[0] in ('Function Name',):
        continue
    try:
        line = int(r[0]); s = int(r[hdr['# Samples']])
    except (ValueError, IndexError):
        continue
    if r[2] != '-':        # a SASS row under a CUDA line: skip (the CUDA row carries the aggregate)
        continue
    tot += s
    key = (cur_file.split('/')[-1], line)
    a = agg.setdefault(key, [0, r[1][:110]])
    a[0] += s
print('total samples', tot)
for (f, line), (s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    if want in f:
        print(f'{s:7d} {100*s/max(tot,1):5.1f}%  {f}:{line:<5d} {src}')
