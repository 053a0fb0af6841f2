#!/bin/bash
# GPU session r2l (1 GPU): null handling + codecs in the default suite, rows kernel with direct field loads (A/B), the full default bench line
mkdir -p gpurun_out
echo "== tests default"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2l_tests_default.log 2>&1; tail -4 gpurun_out/r2l_tests_default.log
echo "== tests PB_AGG_ROWS_DIRECT=1, smem always"; PB_AGG_ROWS_DIRECT=1 PB_AGG_SMEM_MIN=0 timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_tests_direct.log 2>&1; tail -3 gpurun_out/r2l_tests_direct.log
echo "== tests PB_AGG_ROWS_DIRECT=1"; PB_AGG_ROWS_DIRECT=1 timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r2l_tests_direct2.log 2>&1; tail -3 gpurun_out/r2l_tests_direct2.log
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-variants"
echo "== bench direct"; PB_AGG_ROWS_DIRECT=1 timeout 300 $B > gpurun_out/r2l_bench_direct.json 2> gpurun_out/r2l_bench_direct.err; tail -c 300 gpurun_out/r2l_bench_direct.err
echo "== bench sel25 direct"; PB_AGG_ROWS_DIRECT=1 timeout 300 $B --in-values 500 > gpurun_out/r2l_bench_sel25_direct.json 2> gpurun_out/r2l_bench_sel25_direct.err; tail -c 300 gpurun_out/r2l_bench_sel25_direct.err
echo "== bench sel25 default"; timeout 300 $B --in-values 500 > gpurun_out/r2l_bench_sel25.json 2> gpurun_out/r2l_bench_sel25.err; tail -c 300 gpurun_out/r2l_bench_sel25.err
echo "== bench default, full line (e2e, cpu baselines, variants)"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench_full.json 2> gpurun_out/r2l_bench_full.err; tail -c 300 gpurun_out/r2l_bench_full.err
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_ref.json 2> gpurun_out/r2l_ref.err; tail -c 400 gpurun_out/r2l_ref.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2l_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value %.4g ms/step %.4f filter %.4f agg %.4f dev %.4f launches %s parity %s" % (d["value"], d["ms_per_step"], d["filter_kernel_ms"], d["agg_kernel_ms"], d["device_ms_per_step"], d["gpu_launches"], d.get("parity_checked")))
        if d.get("cpu_baseline"): print("  cpu8", d["cpu_baseline"]["value"], "all", (d.get("cpu_baseline_all_cores") or {}).get("value"), "e2e", d["e2e"]["value"])
        s = d.get("selectivity_25pct")
        if s: print("  sel25", s["ms_per_step"], s["agg_kernel_ms"], s["whole_query_frac_on_step_time"])
    except Exception as e:
        print(f, "ERR", e)
PY
