#!/bin/bash
# GPU session r2k (8 GPUs): 4-rank parity test, bench at N = 8 / 4 (weak + strong + sel25, parity-checked), BASELINE configs 3 and 4 on 8 GPUs, 5 on 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== multi-GPU test, 4 ranks"; timeout 300 python -m pytest tests/test_gpu_multi.py -q -k "ranks_merge and 4" > gpurun_out/r2k_tests_multi.log 2>&1; tail -3 gpurun_out/r2k_tests_multi.log
for N in 8; do
  echo "== bench N=$N"; timeout 200 $T --nproc-per-node $N --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2k_bench_n$N.json 2> gpurun_out/r2k_bench_n$N.err; echo "rc=$?"; tail -c 300 gpurun_out/r2k_bench_n$N.err
done
echo "== configs 3,4 on 8 GPUs (scale ${CFG_SCALE:-0.25})"; timeout 240 $T --nproc-per-node 8 --master-port 29521 bench_configs.py --only 3,4 --scale ${CFG_SCALE:-0.25} --steps 5 > gpurun_out/r2k_configs_n8.jsonl 2> gpurun_out/r2k_configs_n8.err; echo "rc=$?"; tail -c 300 gpurun_out/r2k_configs_n8.err
echo "== config 5 on 4 GPUs (scale ${CFG_SCALE:-0.25})"; timeout 200 $T --nproc-per-node 4 --master-port 29522 bench_configs.py --only 5 --scale ${CFG_SCALE:-0.25} --steps 5 > gpurun_out/r2k_configs_n4.jsonl 2> gpurun_out/r2k_configs_n4.err; echo "rc=$?"; tail -c 300 gpurun_out/r2k_configs_n4.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2k_bench_n*.json")):
    try:
        d = json.load(open(f))
        print("N", d["n_gpus"], "value %.4g ms/step %.4f filter %.4f agg %.4f nccl %.4f parity %s e2e %.4g" % (d["value"], d["ms_per_step"], d["filter_kernel_ms"], d["agg_kernel_ms"], d.get("nccl_merge_ms") or 0, d.get("parity_checked"), (d.get("e2e") or {}).get("value") or 0))
        print("  strong", d["strong"]["ms_per_step"], d["strong"]["value"], d["strong"]["breakdown_ms"])
        s = d.get("selectivity_25pct"); print("  sel25", s["ms_per_step"], s["agg_kernel_ms"], s.get("parity_checked"))
    except Exception as e:
        print(f, "ERR", e)
for f in sorted(glob.glob("gpurun_out/r2k_configs_n*.jsonl")):
    for l in open(f):
        try:
            d = json.loads(l); print("config", d["config"], "N", d["n_gpus"], "rows/s %.4g wall %.3f ms kernels %.3f nccl %.3f parity %s" % (d["rows_per_s_wall"], d["wall_ms_per_step"], d["kernel_ms"], d["nccl_merge_ms"], d["parity_vs_oracle"][:70]))
        except Exception as e:
            print("ERR", e, l[:200])
PY
