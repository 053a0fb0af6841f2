#!/bin/bash
# GPU session r2h (1 GPU): shuffle-LUT filter kernel + candidate batching, exact-integer shared-memory sums + 4 rows in flight,
# chunk-compressed raw columns, A/B of each
mkdir -p gpurun_out
echo "== tests default"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2h_tests_default.log 2>&1; tail -4 gpurun_out/r2h_tests_default.log
echo "== tests no shfl / exact-int off / smem always"; PB_FILTER_SHFL=0 PB_AGG_EXACT_INT=0 PB_AGG_SMEM_MIN=0 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_tests_alt.log 2>&1; tail -3 gpurun_out/r2h_tests_alt.log
echo "== tests smem always (exact-int on), no graph"; PB_AGG_SMEM_MIN=0 PB_GRAPH=0 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2h_tests_smem.log 2>&1; tail -3 gpurun_out/r2h_tests_smem.log
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-variants"
echo "== bench default"; timeout 400 $B > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 300 gpurun_out/r2h_bench.err
echo "== bench PB_FILTER_SHFL=0"; PB_FILTER_SHFL=0 timeout 400 $B > gpurun_out/r2h_bench_noshfl.json 2> gpurun_out/r2h_bench_noshfl.err; tail -c 300 gpurun_out/r2h_bench_noshfl.err
echo "== bench sel25"; timeout 400 $B --in-values 500 > gpurun_out/r2h_bench_sel25.json 2> gpurun_out/r2h_bench_sel25.err; tail -c 300 gpurun_out/r2h_bench_sel25.err
echo "== bench sel25 PB_AGG_EXACT_INT=0"; PB_AGG_EXACT_INT=0 timeout 400 $B --in-values 500 > gpurun_out/r2h_bench_sel25_noexact.json 2> gpurun_out/r2h_bench_sel25_noexact.err; tail -c 300 gpurun_out/r2h_bench_sel25_noexact.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2h_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value %.4g ms/step %.4f filter %.4f agg %.4f dev %.4f host_us %s launches %s parity %s" % (d["value"], d["ms_per_step"], d["filter_kernel_ms"], d["agg_kernel_ms"], d["device_ms_per_step"], d["host_us_by_phase"], d["gpu_launches"], d.get("parity_checked")))
    except Exception as e:
        print(f, "ERR", e)
PY
echo "== ncu full: filter + agg kernels of cached plans"
PB_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pb_filter" -s 12 -c 1 -o gpurun_out/r2h_prof_filter python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants > gpurun_out/r2h_ncu_filter.out 2>&1; tail -2 gpurun_out/r2h_ncu_filter.out | cut -c1-200
PB_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pb_agg" -s 12 -c 1 -o gpurun_out/r2h_prof_agg_sel25 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-variants --in-values 500 > gpurun_out/r2h_ncu_agg25.out 2>&1; tail -2 gpurun_out/r2h_ncu_agg25.out | cut -c1-200
ls -la gpurun_out/r2h*.ncu-rep
