#!/bin/bash
# GPU session r2b: parity suite under the default plan and with every new path forced, then the bench A/B
mkdir -p gpurun_out
nvidia-smi -L | head -3
echo "== default"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests_default.log 2>&1; tail -3 gpurun_out/r2b_tests_default.log
echo "== fuse always"; PB_FUSE_PERMILLE=1000 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests_fuse.log 2>&1; tail -3 gpurun_out/r2b_tests_fuse.log
echo "== smem always, never fuse"; PB_FUSE_PERMILLE=-1 PB_AGG_SMEM_MIN=0 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests_smem.log 2>&1; tail -3 gpurun_out/r2b_tests_smem.log
echo "== bench default"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 300 gpurun_out/r2b_bench.err
echo "== bench nofuse nosmem"; PB_FUSE_PERMILLE=-1 PB_AGG_SMEM=0 timeout 900 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/r2b_bench_old.json 2> gpurun_out/r2b_bench_old.err; tail -c 300 gpurun_out/r2b_bench_old.err
python - <<'PY'
import json
for f in ("r2b_bench", "r2b_bench_old"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, "value %.4g ms/step %.4f filter %.4f agg %.4f dev %.4f host_us %s" % (d["value"], d["ms_per_step"], d["filter_kernel_ms"], d["agg_kernel_ms"], d["device_ms_per_step"], d["host_us_by_phase"]))
        s = d.get("selectivity_25pct"); print("  sel25", s and {k: s[k] for k in ("ms_per_step", "filter_kernel_ms", "agg_kernel_ms", "whole_query_frac_on_step_time")})
        o = d.get("strong"); print("  strong", o and {k: o[k] for k in ("ms_per_step", "value", "breakdown_ms")})
    except Exception as e:
        print(f, "ERR", e)
PY
