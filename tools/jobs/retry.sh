#!/bin/bash
# usage: tools/jobs/retry.sh <gpus> <timeout> <out-file> <command...>   — retries gpurun while the pod has no free slot
gpus=$1; to=$2; out=$3; shift 3
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --gpus "$gpus" --timeout "$to" -- "$@" > "$out" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "[retry] finished rc=$rc after $i attempt(s)" >> "$out"; exit $rc; fi
  sleep 90
done
echo "[retry] gave up" >> "$out"; exit 3
