#!/usr/bin/env bash
# B200-native launch: one rank per GPU, servers live in HBM shards, push/pull are fused in-kernel NVSwitch collectives.
# usage: run_fabric.sh [NGPU=8] [PARTIES=2] [example=cnn.py] [script args...]
NGPU=${1:-8}; PARTIES=${2:-2}; EX=${3:-cnn.py}; shift 3 || true
HERE=$(cd "$(dirname "$0")" && pwd)
GEOMX_NUM_PARTIES=$PARTIES exec python -m torch.distributed.run --nnodes=1 --nproc-per-node "$NGPU" --master-addr 127.0.0.1 --master-port ${MASTER_PORT:-29400} \
  "$HERE/../../examples/$EX" "$@"
