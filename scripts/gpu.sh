#!/bin/bash
# Rebuild the in-tree .so, then run a command on the B200 box: scripts/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python stable-diffusion_b200/build.py >/dev/null
T=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
