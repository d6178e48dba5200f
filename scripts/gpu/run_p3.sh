#!/usr/bin/env bash
# P3: priority-ordered push slices, the push response carries the parameters.
# Reference counterpart: scripts/gpu/run_p3.sh (12 local processes; differences vs vanilla are the env vars / script below).
HERE=$(cd "$(dirname "$0")" && pwd)
EXTRA_SERVER_ENV="ENABLE_P3=1" EXTRA_WORKER_ENV="ENABLE_P3=1" MASTER_ARGS="" N_GS=1 \
  exec "$HERE/../hips_launch.sh" gpu "$HERE/../../examples/cnn.py"  "$@"
