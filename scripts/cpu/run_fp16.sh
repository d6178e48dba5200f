#!/usr/bin/env bash
# FP16 transport: tensors cross the tiers as float16.
# Reference counterpart: scripts/cpu/run_fp16.sh (12 local processes; differences vs vanilla are the env vars / script below).
HERE=$(cd "$(dirname "$0")" && pwd)
EXTRA_SERVER_ENV="" EXTRA_WORKER_ENV="" MASTER_ARGS="" N_GS=1 \
  exec "$HERE/../hips_launch.sh" cpu "$HERE/../../examples/cnn_fp16.py"  "$@"
