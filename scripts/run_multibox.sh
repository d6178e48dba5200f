#!/usr/bin/env bash
# Several NVSwitch boxes, one HiPS job: every box is a party (torchrun, one rank per GPU, NVLink collectives inside the box), only the box
# leader talks to the global server over TCP (geomx_b200/kvstore/hybrid.py).  This script emulates BOXES boxes on ONE host for a smoke run;
# on real clusters run the scheduler / server block on the central site and one "box" block per machine with DMLC_PS_ROOT_URI pointing there.
# usage: run_multibox.sh <cpu|gpu> <example.py> [script args...]     env knobs: BOXES (2) RANKS_PER_BOX (2) BASE_PORT (9392) LOG_DIR
set -euo pipefail
MODE=${1:?cpu|gpu}; SCRIPT=${2:?example script}; shift 2
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(dirname "$HERE")
PY=${PYTHON:-python}
BOXES=${BOXES:-2}; RPB=${RANKS_PER_BOX:-2}; BASE_PORT=${BASE_PORT:-9392}
LOG_DIR=${LOG_DIR:-/tmp/geomx_multibox}; mkdir -p "$LOG_DIR"
CPU_FLAG=""; [ "$MODE" = "cpu" ] && CPU_FLAG="--cpu"
PS="DMLC_PS_ROOT_URI=127.0.0.1 DMLC_PS_ROOT_PORT=$BASE_PORT DMLC_NUM_SERVER=1 DMLC_NUM_WORKER=$BOXES DMLC_NUM_ALL_WORKER=$((BOXES * RPB))"
COMMON="GEOMX_SYNTHETIC_SIZE=${GEOMX_SYNTHETIC_SIZE:-4096} PYTHONPATH=$ROOT"
pids=()
env $COMMON $PS DMLC_ROLE=scheduler $PY -c "import geomx_b200" > "$LOG_DIR/scheduler.log" 2>&1 & pids+=($!)
env $COMMON $PS DMLC_ROLE=server $PY -c "import geomx_b200" > "$LOG_DIR/server.log" 2>&1 & pids+=($!)
slice=0
for b in $(seq 1 "$BOXES"); do
  MPORT=$((BASE_PORT + 10 + b))
  for r in $(seq 0 $((RPB - 1))); do
    DEV="LOCAL_RANK=$r"; [ "$MODE" = "gpu" ] && DEV="LOCAL_RANK=$slice"
    env $COMMON $PS DMLC_ROLE=worker RANK=$r WORLD_SIZE=$RPB MASTER_ADDR=127.0.0.1 MASTER_PORT=$MPORT $DEV \
      $PY "$SCRIPT" $CPU_FLAG --data-slice-idx $slice "$@" > "$LOG_DIR/box${b}_rank${r}.log" 2>&1 & pids+=($!)
    slice=$((slice + 1))
  done
done
echo "launched ${#pids[@]} processes ($BOXES boxes x $RPB ranks + scheduler + server); logs in $LOG_DIR"
trap 'kill "${pids[@]}" 2>/dev/null' EXIT INT TERM
rc=0
for pid in "${pids[@]}"; do r=0; wait "$pid" || r=$?; [ $r -ne 0 ] && [ $rc -eq 0 ] && rc=$r; done
trap - EXIT
tail -n 3 "$LOG_DIR/box1_rank0.log"
exit $rc
