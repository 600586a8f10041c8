"""Extracts the judged metrics from an .ncu-rep into a small CSV:  python tools/ncu_summary.py in.ncu-rep out.csv"""
import csv
import subprocess
import sys

PREFIXES = ('gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'gpu__dram_throughput.avg.pct', 'lts__throughput.avg.pct', 'l1tex__throughput.avg.pct', 'sm__throughput.avg.pct',
            'launch__registers_per_thread', 'sm__warps_active.avg.pct', 'sm__cycles_elapsed.avg', 'launch__shared_mem_per_block_dynamic',
            'smsp__inst_executed.sum', 'launch__grid_size', 'launch__block_size', 'dram__cycles_active.avg.pct')
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = [i for i, h in enumerate(hdr) if h in ('ID', 'Kernel Name') or any(h.startswith(p) for p in PREFIXES)]
with open(sys.argv[2], 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow([hdr[i] for i in keep])
    w.writerow([units[i] for i in keep])
    for r in rows[2:]:
        w.writerow([r[i] for i in keep])
