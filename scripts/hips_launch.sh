#!/usr/bin/env bash
# Launch the reference's demo topology on ONE host over the native TCP HiPS transport:
#   global scheduler + N_GS global server(s) + master worker + central scheduler + PARTIES x (scheduler, server, WPP workers)
# (default 1 + 1 + 1 + 1 + 2 x (1 + 1 + 2) = 12 processes, exactly scripts/{cpu,gpu}/run_vanilla_hips.sh of the reference).
# usage: hips_launch.sh <cpu|gpu> <example.py> [script args...]      env knobs: PARTIES WPP N_GS BASE_PORT EXTRA_SERVER_ENV EXTRA_WORKER_ENV LOG_DIR
set -euo pipefail
MODE=${1:?cpu|gpu}; SCRIPT=${2:?example script}; shift 2
HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(dirname "$HERE")
PY=${PYTHON:-python}
PARTIES=${PARTIES:-2}; WPP=${WPP:-2}; N_GS=${N_GS:-1}; BASE_PORT=${BASE_PORT:-9092}
LOG_DIR=${LOG_DIR:-/tmp/geomx_logs}; mkdir -p "$LOG_DIR"
ALLW=$((PARTIES * WPP))
CPU_FLAG=""; [ "$MODE" = "cpu" ] && CPU_FLAG="--cpu"
BOOT="import sys; sys.path.insert(0, '$ROOT'); import geomx_b200"
GLOBAL_ENV="DMLC_PS_GLOBAL_ROOT_URI=127.0.0.1 DMLC_PS_GLOBAL_ROOT_PORT=$BASE_PORT DMLC_NUM_GLOBAL_SERVER=$N_GS DMLC_NUM_GLOBAL_WORKER=$PARTIES"
COMMON="PS_VERBOSE=${PS_VERBOSE:-0} GEOMX_SYNTHETIC_SIZE=${GEOMX_SYNTHETIC_SIZE:-4096}"
pids=()
run() { env $COMMON "$@" & pids+=($!); }

# --- central party ---------------------------------------------------------------------------------------------------------------
run $GLOBAL_ENV DMLC_ROLE_GLOBAL=global_scheduler ${EXTRA_SERVER_ENV:-} $PY -c "$BOOT" > "$LOG_DIR/global_scheduler.log" 2>&1
CPORT=$((BASE_PORT + 1))
CENTRAL="DMLC_PS_ROOT_URI=127.0.0.1 DMLC_PS_ROOT_PORT=$CPORT DMLC_NUM_SERVER=$N_GS DMLC_NUM_WORKER=1 DMLC_NUM_ALL_WORKER=$ALLW"
for g in $(seq 1 "$N_GS"); do
  run $GLOBAL_ENV $CENTRAL DMLC_ROLE_GLOBAL=global_server DMLC_ROLE=server DMLC_ENABLE_CENTRAL_WORKER=0 ${EXTRA_SERVER_ENV:-} $PY -c "$BOOT" > "$LOG_DIR/global_server$g.log" 2>&1
done
run $CENTRAL DMLC_ROLE=scheduler ${EXTRA_SERVER_ENV:-} $PY -c "$BOOT" > "$LOG_DIR/central_scheduler.log" 2>&1
run $CENTRAL DMLC_ROLE=worker DMLC_ROLE_MASTER_WORKER=1 ${EXTRA_WORKER_ENV:-} ${MASTER_WORKER_ENV:-} $PY "$SCRIPT" $CPU_FLAG ${MASTER_ARGS:-} "$@" > "$LOG_DIR/master_worker.log" 2>&1

# --- participating parties ---------------------------------------------------------------------------------------------------------
slice=0; last=""
for p in $(seq 1 "$PARTIES"); do
  PORT=$((BASE_PORT + 1 + p))
  PARTY="DMLC_PS_ROOT_URI=127.0.0.1 DMLC_PS_ROOT_PORT=$PORT DMLC_NUM_SERVER=1 DMLC_NUM_WORKER=$WPP DMLC_NUM_ALL_WORKER=$ALLW"
  run $PARTY DMLC_ROLE=scheduler ${EXTRA_SERVER_ENV:-} $PY -c "$BOOT" > "$LOG_DIR/party${p}_scheduler.log" 2>&1
  run $GLOBAL_ENV $PARTY DMLC_ROLE=server ${EXTRA_SERVER_ENV:-} $PY -c "$BOOT" > "$LOG_DIR/party${p}_server.log" 2>&1
  for w in $(seq 1 "$WPP"); do
    DEV=""; [ "$MODE" = "gpu" ] && DEV="LOCAL_RANK=$slice"
    run $PARTY DMLC_ROLE=worker $DEV ${EXTRA_WORKER_ENV:-} $PY "$SCRIPT" $CPU_FLAG --data-slice-idx $slice "$@" > "$LOG_DIR/party${p}_worker${w}.log" 2>&1
    last="$LOG_DIR/party${p}_worker${w}.log"; slice=$((slice + 1))
  done
done
echo "launched ${#pids[@]} processes; logs in $LOG_DIR (tail -f $last)"
# Fail fast: a role that dies at start-up (port in use, bad environment ...) would otherwise leave every other process waiting at the
# rendezvous forever.  Poll the children; on the first non-zero exit stop the whole job and report which log to read.
trap 'kill "${pids[@]}" 2>/dev/null' EXIT INT TERM
rc=0; alive=("${pids[@]}")
while [ ${#alive[@]} -gt 0 ]; do
  next=()
  for pid in "${alive[@]}"; do
    if kill -0 "$pid" 2>/dev/null; then
      next+=("$pid")
    else
      r=0; wait "$pid" || r=$?
      if [ $r -ne 0 ] && [ $rc -eq 0 ]; then rc=$r; echo "process $pid exited with code $r - stopping the job (see $LOG_DIR)" >&2; fi
    fi
  done
  alive=("${next[@]}")
  if [ $rc -ne 0 ]; then
    for pid in "${alive[@]}"; do kill "$pid" 2>/dev/null; done
    sleep 1
    for pid in "${alive[@]}"; do kill -9 "$pid" 2>/dev/null; done
    break
  fi
  [ ${#alive[@]} -gt 0 ] && sleep 0.3
done
trap - EXIT
tail -n 3 "$last"
exit $rc
