#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> <gpus> '<command>'   -- retries while the pod is busy (nothing is charged for those)
T=$1; G=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then OUT=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1); else OUT=$(/usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@" 2>&1); fi
  if echo "$OUT" | grep -q "status=transient"; then echo "[retry $i] busy"; sleep 90; continue; fi
  echo "$OUT" | cut -c1-1500 | tail -80
  exit 0
done
echo "gave up"
