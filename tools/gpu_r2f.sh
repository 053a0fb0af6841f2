#!/bin/bash
# GPU session r2f: pb_agg_rows_kernel, spec filter at 4 CTAs/SM, fixes (raw DC, trim test)
mkdir -p gpurun_out
echo "== tests default"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests_default.log 2>&1; tail -4 gpurun_out/r2f_tests_default.log
echo "== tests no rows kernel, no row groups, no plan cache"; PB_AGG_ROWS=0 PB_ROW_GROUPS=0 PB_PLAN_CACHE=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests_norg.log 2>&1; tail -3 gpurun_out/r2f_tests_norg.log
echo "== tests no graph, smem always, no spec"; PB_GRAPH=0 PB_AGG_SMEM_MIN=0 PB_FILTER_SPEC=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2f_tests_nograph.log 2>&1; tail -3 gpurun_out/r2f_tests_nograph.log
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
echo "== bench default (full)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; tail -c 300 gpurun_out/r2f_bench.err
for v in "PB_AGG_ROWS=0" "PB_AGG_SMEM_MIN=0" "PB_AGG_SMEM=0"; do
  n=$(echo "$v" | sed 's/[^A-Za-z0-9]/_/g')
  echo "== bench $v"; env $v timeout 600 $B --no-variants > gpurun_out/r2f_bench_$n.json 2> gpurun_out/r2f_bench_$n.err; tail -c 200 gpurun_out/r2f_bench_$n.err
done
echo "== bench sel25 variants"
timeout 600 $B --in-values 500 --no-variants > gpurun_out/r2f_bench_sel25.json 2> gpurun_out/r2f_bench_sel25.err
PB_AGG_ROWS=0 timeout 600 $B --in-values 500 --no-variants > gpurun_out/r2f_bench_sel25_norows.json 2> gpurun_out/r2f_bench_sel25_norows.err
PB_AGG_SMEM=0 timeout 600 $B --in-values 500 --no-variants > gpurun_out/r2f_bench_sel25_nosmem.json 2> gpurun_out/r2f_bench_sel25_nosmem.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2f_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value %.4g ms/step %.4f filter %.4f agg %.4f dev %.4f host_us %s launches %s" % (d["value"], d["ms_per_step"], d["filter_kernel_ms"], d["agg_kernel_ms"], d["device_ms_per_step"], d["host_us_by_phase"], d["gpu_launches"]))
        s = d.get("selectivity_25pct")
        if s: print("  sel25", {k: s[k] for k in ("ms_per_step", "filter_kernel_ms", "agg_kernel_ms", "whole_query_frac_on_step_time")})
        o = d.get("strong")
        if o: print("  strong", {k: o[k] for k in ("ms_per_step", "value", "breakdown_ms")})
    except Exception as e:
        print(f, "ERR", e)
PY
echo "== ncu full: agg kernels (headline + sel25), cached plans only (-s skips the cold waves)"
PB_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pb_agg" -s 12 -c 2 -o gpurun_out/r2f_prof_agg python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/r2f_ncu_agg.out 2>&1; tail -2 gpurun_out/r2f_ncu_agg.out | cut -c1-200
PB_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pb_agg" -s 12 -c 1 -o gpurun_out/r2f_prof_agg_sel25 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --in-values 500 > gpurun_out/r2f_ncu_agg25.out 2>&1; tail -2 gpurun_out/r2f_ncu_agg25.out | cut -c1-200
PB_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pb_filter" -s 12 -c 1 -o gpurun_out/r2f_prof_filter python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/r2f_ncu_filter.out 2>&1; tail -2 gpurun_out/r2f_ncu_filter.out | cut -c1-200
ls -la gpurun_out/r2f*.ncu-rep
