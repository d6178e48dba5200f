#!/usr/bin/env bash
# TSEngine: scheduler-driven merge / relay overlay (intra- and inter-party).
# Reference counterpart: scripts/cpu/run_tsengine.sh (12 local processes; differences vs vanilla are the env vars / script below).
HERE=$(cd "$(dirname "$0")" && pwd)
EXTRA_SERVER_ENV="ENABLE_INTER_TS=1 ENABLE_INTRA_TS=1 MAX_GREED_RATE_TS=0.9" EXTRA_WORKER_ENV="ENABLE_INTER_TS=1 ENABLE_INTRA_TS=1 MAX_GREED_RATE_TS=0.9" MASTER_ARGS="" N_GS=1 \
  exec "$HERE/../hips_launch.sh" cpu "$HERE/../../examples/cnn.py"  "$@"
