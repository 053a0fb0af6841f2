#!/bin/bash
# GPU session r2g (2 GPUs): N > 1 parity tests, bench at N = 1 and 2, bench_configs under torchrun
mkdir -p gpurun_out
nvidia-smi -L
echo "== tests default (2 GPUs visible)"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_tests_default.log 2>&1; tail -6 gpurun_out/r2g_tests_default.log
B="python bench.py --steps 20 --warmup 5"
echo "== bench N=1"; timeout 600 $B --no-e2e --no-cpu-baseline --no-variants > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; tail -c 200 gpurun_out/r2g_bench_n1.err
echo "== bench N=1 sel25"; timeout 600 $B --no-e2e --no-cpu-baseline --no-variants --in-values 500 > gpurun_out/r2g_bench_n1_sel25.json 2> gpurun_out/r2g_bench_n1_sel25.err; tail -c 200 gpurun_out/r2g_bench_n1_sel25.err
echo "== bench N=2 (torchrun)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; tail -c 600 gpurun_out/r2g_bench_n2.err
echo "== bench_configs N=2 (scaled)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench_configs.py --only 3,4,5 --scale 0.1 --steps 5 > gpurun_out/r2g_configs_n2.jsonl 2> gpurun_out/r2g_configs_n2.err; tail -c 600 gpurun_out/r2g_configs_n2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2g_bench*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "N", d["n_gpus"], "value %.4g ms/step %.4f filter %.4f agg %.4f nccl %.4f dev %.4f host_us %s parity %s" % (d["value"], d["ms_per_step"], d["filter_kernel_ms"], d["agg_kernel_ms"], d.get("nccl_merge_ms") or 0, d["device_ms_per_step"], d["host_us_by_phase"], d.get("parity_checked")))
        s = d.get("selectivity_25pct")
        if s: print("  sel25", {k: s[k] for k in ("ms_per_step", "filter_kernel_ms", "agg_kernel_ms", "whole_query_frac_on_step_time")})
        o = d.get("strong")
        if o: print("  strong", {k: o[k] for k in ("ms_per_step", "value", "breakdown_ms")})
        e = d.get("e2e")
        if e: print("  e2e", e["value"], e["ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
for l in open("gpurun_out/r2g_configs_n2.jsonl"):
    try:
        d = json.loads(l); print("config", d["config"], "N", d["n_gpus"], "rows/s %.4g wall %.3f ms kernels %.3f nccl %.3f parity %s" % (d["rows_per_s_wall"], d["wall_ms_per_step"], d["kernel_ms"], d["nccl_merge_ms"], d["parity_vs_oracle"][:90]))
    except Exception as e:
        print("ERR", e, l[:200])
PY
