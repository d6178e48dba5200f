#!/usr/bin/env bash
# MultiGPS: 2 global servers; big tensors partitioned, small ones hashed.
# Reference counterpart: scripts/gpu/run_multi_gps.sh (12 local processes; differences vs vanilla are the env vars / script below).
HERE=$(cd "$(dirname "$0")" && pwd)
EXTRA_SERVER_ENV="MXNET_KVSTORE_BIGARRAY_BOUND=10000" EXTRA_WORKER_ENV="MXNET_KVSTORE_BIGARRAY_BOUND=10000" MASTER_ARGS="" N_GS=2 \
  exec "$HERE/../hips_launch.sh" gpu "$HERE/../../examples/cnn.py"  "$@"
