#!/bin/bash
# GPU session r2d: row groups A/B, fused with row groups, graph replay
mkdir -p gpurun_out
echo "== tests default"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests_default.log 2>&1; tail -3 gpurun_out/r2d_tests_default.log
echo "== tests row groups off, fuse always"; PB_ROW_GROUPS=0 PB_FUSE_PERMILLE=1000 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests_norg.log 2>&1; tail -3 gpurun_out/r2d_tests_norg.log
echo "== tests fuse always + row groups, no graph"; PB_GRAPH=0 PB_FUSE_PERMILLE=1000 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_tests_fuse.log 2>&1; tail -3 gpurun_out/r2d_tests_fuse.log
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline"
echo "== bench default (full)"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; tail -c 300 gpurun_out/r2d_bench.err
for v in "PB_ROW_GROUPS=0" "PB_GRAPH=0" "PB_FUSE_PERMILLE=50" "PB_FUSE_PERMILLE=50 PB_FUSE_BATCH=128" "PB_FILTER_SPEC=0" "PB_FILTER_SPEC=0 PB_ROW_GROUPS=0"; do
  n=$(echo "$v" | sed 's/[^A-Za-z0-9]/_/g')
  echo "== bench $v"; env $v timeout 600 $B --no-variants > gpurun_out/r2d_bench_$n.json 2> gpurun_out/r2d_bench_$n.err; tail -c 200 gpurun_out/r2d_bench_$n.err
done
echo "== bench sel25 variants"; PB_ROW_GROUPS=0 timeout 600 $B --in-values 500 --no-variants > gpurun_out/r2d_bench_sel25_norg.json 2> gpurun_out/r2d_bench_sel25_norg.err
timeout 600 $B --in-values 500 --no-variants > gpurun_out/r2d_bench_sel25.json 2> gpurun_out/r2d_bench_sel25.err
PB_AGG_SMEM=0 timeout 600 $B --in-values 500 --no-variants > gpurun_out/r2d_bench_sel25_nosmem.json 2> gpurun_out/r2d_bench_sel25_nosmem.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d_bench*.json")):
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
echo "== bench_configs small scale"; date; timeout 600 python bench_configs.py --only 1,3,4,5 --scale 0.05 --steps 3 > gpurun_out/r2d_configs_small.jsonl 2> gpurun_out/r2d_configs_small.err; tail -c 400 gpurun_out/r2d_configs_small.err; cut -c1-700 gpurun_out/r2d_configs_small.jsonl
