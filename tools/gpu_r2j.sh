#!/bin/bash
# GPU session r2j (2 GPUs): does a torchrun bench exit cleanly now that comm-captured graphs are destroyed before the communicator?
mkdir -p gpurun_out
nvidia-smi -L
echo "== multi-GPU tests"; timeout 600 python -m pytest tests/test_gpu_multi.py -q > gpurun_out/r2j_tests_multi.log 2>&1; tail -3 gpurun_out/r2j_tests_multi.log
echo "== null-handling + filtered-aggregation tests"; timeout 400 python -m pytest tests/test_gpu_null_handling.py tests/test_gpu_parity.py -q -k "null or filtered or FILTER or golden" > gpurun_out/r2j_tests_null.log 2>&1; tail -3 gpurun_out/r2j_tests_null.log
echo "== bench N=2 (torchrun)"; /usr/bin/time -f "wall %e s" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err; echo "rc=$?"; tail -c 400 gpurun_out/r2j_bench_n2.err
echo "== bench N=2 reference arm"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2j_ref_n2.json 2> gpurun_out/r2j_ref_n2.err; echo "rc=$?"; tail -c 300 gpurun_out/r2j_ref_n2.json
echo "== bench_configs N=2 (scaled)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench_configs.py --only 3,4,5 --scale 0.1 --steps 5 > gpurun_out/r2j_configs_n2.jsonl 2> gpurun_out/r2j_configs_n2.err; echo "rc=$?"; tail -c 300 gpurun_out/r2j_configs_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2j_bench_n2.json"))
    print("N", d["n_gpus"], "value %.4g ms/step %.4f filter %.4f agg %.4f nccl %.4f parity %s" % (d["value"], d["ms_per_step"], d["filter_kernel_ms"], d["agg_kernel_ms"], d.get("nccl_merge_ms") or 0, d.get("parity_checked")))
    print("  strong", d["strong"]["ms_per_step"], d["strong"]["breakdown_ms"])
except Exception as e:
    print("ERR", e)
for l in open("gpurun_out/r2j_configs_n2.jsonl"):
    try:
        d = json.loads(l); print("config", d["config"], "N", d["n_gpus"], "rows/s %.4g wall %.3f ms parity %s" % (d["rows_per_s_wall"], d["wall_ms_per_step"], d["parity_vs_oracle"][:60]))
    except Exception as e:
        print("ERR", e, l[:200])
PY
