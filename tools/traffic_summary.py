"""ncu metric CSV of one profiled step (tools/ncu_gemm_traffic.sh) -> profiles/gemm_traffic.json, the DRAM bytes per
tcgen05 GEMM launch that bench.py quotes as `roofline.traffic`:
    python tools/traffic_summary.py gpurun_out/r2_gemm_traffic.csv profiles/r2_gemm_traffic.csv"""
import csv
import datetime
import json
import os
import shutil
import sys

src, kept = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
per = {}
for r in rows:
    per.setdefault(int(r[0]), {"kernel": r[4]})[r[12]] = float(r[14].replace(",", ""))
gemm = [v for v in per.values() if "gemm_tc_kernel" in v["kernel"]]
rd = sum(v.get("dram__bytes_read.sum", 0.0) for v in gemm)
wr = sum(v.get("dram__bytes_write.sum", 0.0) for v in gemm)
n = len(gemm)
if os.path.abspath(src) != os.path.abspath(kept):
    shutil.copyfile(src, kept)
out = {"source": kept, "when": datetime.date.today().isoformat(), "launches": n, "bytes_per_launch": round((rd + wr) / n, 1),
       "read_mb_per_launch": rd / n / 1e6, "write_mb_per_launch": wr / n / 1e6,
       "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over the gemm_tc_kernel launches of one "
              "bench.py --profile-step training step (batch 1, bf16)"}
json.dump(out, open(os.path.join(os.path.dirname(kept), "gemm_traffic.json"), "w"), indent=1)
print(out)
