#!/usr/bin/env bash
# DGT: contribution-ranked 4 KiB blocks over prioritised channels (local -> global server).
# Reference counterpart: scripts/cpu/run_dgt.sh (12 local processes; differences vs vanilla are the env vars / script below).
HERE=$(cd "$(dirname "$0")" && pwd)
EXTRA_SERVER_ENV="ENABLE_DGT=2 DMLC_UDP_CHANNEL_NUM=3 DMLC_K=0.8 ADAPTIVE_K_FLAG=1" EXTRA_WORKER_ENV="" MASTER_ARGS="" N_GS=1 \
  exec "$HERE/../hips_launch.sh" cpu "$HERE/../../examples/cnn.py"  "$@"
