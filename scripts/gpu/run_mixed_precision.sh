#!/usr/bin/env bash
# MPQ: large tensors fp32 + Bi-Sparse, small tensors dense fp16.
# Reference counterpart: scripts/gpu/run_mixed_precision.sh (12 local processes; differences vs vanilla are the env vars / script below).
HERE=$(cd "$(dirname "$0")" && pwd)
EXTRA_SERVER_ENV="MXNET_KVSTORE_SIZE_LOWER_BOUND=1000" EXTRA_WORKER_ENV="MXNET_KVSTORE_SIZE_LOWER_BOUND=1000" MASTER_ARGS="" N_GS=1 \
  exec "$HERE/../hips_launch.sh" gpu "$HERE/../../examples/cnn_mpq.py"  "$@"
