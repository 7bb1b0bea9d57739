#!/bin/bash
# scripts/gpu_retry.sh <timeout_s> '<command>' : rebuild, then retry gpurun while the pod answers busy (exit 3 / transient)
cd "$(dirname "$0")/.."
python stable-diffusion_b200/build.py >/dev/null || exit 1
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "gave up: pod busy"; exit 3
