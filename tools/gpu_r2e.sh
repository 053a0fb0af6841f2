#!/bin/bash
# GPU session r2e: decoded row-group fields, min/max read-before-RED, hash-limit repair pass, ORDER BY trim, raw DISTINCTCOUNT; ncu captures
mkdir -p gpurun_out
echo "== tests default"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e_tests_default.log 2>&1; tail -4 gpurun_out/r2e_tests_default.log
echo "== tests no row groups, no plan cache"; PB_ROW_GROUPS=0 PB_PLAN_CACHE=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e_tests_norg.log 2>&1; tail -3 gpurun_out/r2e_tests_norg.log
echo "== tests no graph, smem always, no spec"; PB_GRAPH=0 PB_AGG_SMEM_MIN=0 PB_FILTER_SPEC=0 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2e_tests_nograph.log 2>&1; tail -3 gpurun_out/r2e_tests_nograph.log
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
echo "== bench default (full)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; tail -c 300 gpurun_out/r2e_bench.err
for v in "PB_ROW_GROUPS=0" "PB_FILTER_SPEC=0"; do
  n=$(echo "$v" | sed 's/[^A-Za-z0-9]/_/g')
  echo "== bench $v"; env $v timeout 600 $B --no-variants > gpurun_out/r2e_bench_$n.json 2> gpurun_out/r2e_bench_$n.err; tail -c 200 gpurun_out/r2e_bench_$n.err
done
echo "== bench sel25 variants"
timeout 600 $B --in-values 500 --no-variants > gpurun_out/r2e_bench_sel25.json 2> gpurun_out/r2e_bench_sel25.err
PB_ROW_GROUPS=0 timeout 600 $B --in-values 500 --no-variants > gpurun_out/r2e_bench_sel25_norg.json 2> gpurun_out/r2e_bench_sel25_norg.err
PB_AGG_SMEM=0 timeout 600 $B --in-values 500 --no-variants > gpurun_out/r2e_bench_sel25_nosmem.json 2> gpurun_out/r2e_bench_sel25_nosmem.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2e_bench*.json")):
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
echo "== ncu launch list"; PB_GRAPH=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/r2e_ncu_launches.out 2>&1; tail -2 gpurun_out/r2e_ncu_launches.out | cut -c1-300
echo "== ncu full: filter + agg kernels of the headline query"; PB_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pb_filter_kernel|pb_agg_kernel" -s 8 -c 4 -o gpurun_out/r2e_prof python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/r2e_ncu_full.out 2>&1; tail -2 gpurun_out/r2e_ncu_full.out | cut -c1-300
echo "== ncu full: shared-memory aggregation kernel, 25 % selectivity"; PB_GRAPH=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pb_agg_smem_kernel" -s 2 -c 1 -o gpurun_out/r2e_prof_sel25 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --in-values 500 > gpurun_out/r2e_ncu_sel25.out 2>&1; tail -2 gpurun_out/r2e_ncu_sel25.out | cut -c1-300
ls -la gpurun_out/*.ncu-rep
