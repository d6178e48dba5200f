#!/usr/bin/env bash
# Bi-Sparse compression between local and global servers (tensors above the size bound).
# Reference counterpart: scripts/cpu/run_bisparse_compression.sh (12 local processes; differences vs vanilla are the env vars / script below).
HERE=$(cd "$(dirname "$0")" && pwd)
EXTRA_SERVER_ENV="MXNET_KVSTORE_SIZE_LOWER_BOUND=1000" EXTRA_WORKER_ENV="" MASTER_ARGS="" N_GS=1 \
  exec "$HERE/../hips_launch.sh" cpu "$HERE/../../examples/cnn_bsc.py"  "$@"
