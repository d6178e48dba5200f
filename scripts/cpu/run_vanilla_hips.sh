#!/usr/bin/env bash
# Vanilla HiPS, fully-synchronous algorithm (FSA): dist_sync, Adam on the global server.
# Reference counterpart: scripts/cpu/run_vanilla_hips.sh (12 local processes; differences vs vanilla are the env vars / script below).
HERE=$(cd "$(dirname "$0")" && pwd)
EXTRA_SERVER_ENV="" EXTRA_WORKER_ENV="" MASTER_ARGS="" N_GS=1 \
  exec "$HERE/../hips_launch.sh" cpu "$HERE/../../examples/cnn.py"  "$@"
