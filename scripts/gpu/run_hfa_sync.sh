#!/usr/bin/env bash
# HFA: K1=20 local steps per local sync, K2=10 local syncs per global sync.
# Reference counterpart: scripts/gpu/run_hfa_sync.sh (12 local processes; differences vs vanilla are the env vars / script below).
HERE=$(cd "$(dirname "$0")" && pwd)
EXTRA_SERVER_ENV="MXNET_KVSTORE_USE_HFA=1 MXNET_KVSTORE_HFA_K1=20 MXNET_KVSTORE_HFA_K2=10" EXTRA_WORKER_ENV="MXNET_KVSTORE_USE_HFA=1 MXNET_KVSTORE_HFA_K1=20 MXNET_KVSTORE_HFA_K2=10" MASTER_ARGS="" N_GS=1 \
  exec "$HERE/../hips_launch.sh" gpu "$HERE/../../examples/cnn_hfa.py"  "$@"
