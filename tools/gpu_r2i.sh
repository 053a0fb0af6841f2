#!/bin/bash
# GPU session r2i (1 GPU): back on the r2g kernels (+ exact-integer shared-memory sums), chunk-compressed raw columns,
# null-value vectors, wide raw predicates
mkdir -p gpurun_out
echo "== tests default"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2i_tests_default.log 2>&1; tail -4 gpurun_out/r2i_tests_default.log
echo "== tests smem always, no graph, exact-int off"; PB_AGG_SMEM_MIN=0 PB_GRAPH=0 PB_AGG_EXACT_INT=0 timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2i_tests_alt.log 2>&1; tail -3 gpurun_out/r2i_tests_alt.log
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-variants"
echo "== bench default"; timeout 400 $B > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 300 gpurun_out/r2i_bench.err
echo "== bench sel25"; timeout 400 $B --in-values 500 > gpurun_out/r2i_bench_sel25.json 2> gpurun_out/r2i_bench_sel25.err; tail -c 300 gpurun_out/r2i_bench_sel25.err
echo "== bench sel25 PB_AGG_EXACT_INT=0"; PB_AGG_EXACT_INT=0 timeout 400 $B --in-values 500 > gpurun_out/r2i_bench_sel25_noexact.json 2> gpurun_out/r2i_bench_sel25_noexact.err; tail -c 300 gpurun_out/r2i_bench_sel25_noexact.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2i_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "value %.4g ms/step %.4f filter %.4f agg %.4f dev %.4f host_us %s launches %s parity %s" % (d["value"], d["ms_per_step"], d["filter_kernel_ms"], d["agg_kernel_ms"], d["device_ms_per_step"], d["host_us_by_phase"], d["gpu_launches"], d.get("parity_checked")))
    except Exception as e:
        print(f, "ERR", e)
PY
