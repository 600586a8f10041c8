#!/bin/bash
# round-2 final evidence: tests, smoke, both bench arms, attention sweep, launch list, ncu of the tiled attention kernels, config 4
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_final_t_gpu.log 2>&1; echo "pytest gpu exit $?"; tail -2 gpurun_out/r2_final_t_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/r2_final_bench.json
timeout 900 python bench.py --impl reference > gpurun_out/r2_final_bench_ref.json 2> gpurun_out/r2_final_bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/r2_final_bench_ref.json
timeout 900 python bench.py --sweep attn --steps 5 --warmup 2 > gpurun_out/r2_final_attn_sweep.jsonl 2> gpurun_out/r2_final_attn_sweep.err; echo "sweep exit $?"; wc -l gpurun_out/r2_final_attn_sweep.jsonl
timeout 900 python bench.py --workload config4 --batch-per-gpu 2 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2_final_config4_b2.json 2> gpurun_out/r2_final_config4_b2.err; echo "config4 exit $?"; cut -c1-200 gpurun_out/r2_final_config4_b2.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r2_final_launches.csv python bench.py --profile-step --no-cpu-baseline --no-eager-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"; wc -l gpurun_out/r2_final_launches.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_gen" -s 4 -c 4 -f -o gpurun_out/r2_attn_gen python tools/attn_gen_one.py > gpurun_out/ncu_attn_gen.log 2>&1; echo "ncu attn_gen exit $?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"attn_tc_fwd4|attn_tc_bwd3|attn_rowdot3" -s 3 -c 3 -f -o gpurun_out/r2_final_attn python tools/attn_phases.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
