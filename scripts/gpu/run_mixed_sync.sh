#!/usr/bin/env bash
# MixedSync: synchronous inside a party, asynchronous between parties (dist_async on the global tier).
# Reference counterpart: scripts/gpu/run_mixed_sync.sh (12 local processes; differences vs vanilla are the env vars / script below).
HERE=$(cd "$(dirname "$0")" && pwd)
EXTRA_SERVER_ENV="" EXTRA_WORKER_ENV="" MASTER_ARGS="--mixed-sync" N_GS=1 \
  exec "$HERE/../hips_launch.sh" gpu "$HERE/../../examples/cnn.py" --mixed-sync "$@"
