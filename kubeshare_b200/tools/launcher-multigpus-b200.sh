#!/bin/bash
# launcher-multigpus-b200.sh -- drop-in for docker/kubeshare-gemini-scheduler/launcher-multigpus.sh + launcher.py
# (reference launcher-multigpus.sh:21-42, launcher.py:34-80): one gem-arbiter per GPU UUID instead of one gem-schd
# plus an inotify-driven Python supervisor that spawns a gem-pmgr per pod.  No Python, no `inotify` pip package.
#
#   launcher-multigpus-b200.sh <config_dir> <port_dir> [<library_dir>]
#     config_dir   /kubeshare/scheduler/config           (quota file per GPU UUID, written by kubeshare-config)
#     port_dir     /kubeshare/scheduler/podmanagerport   (name/port rows per GPU UUID)
#     library_dir  /kubeshare/library                    (hostPath every pod mounts: pool + quota mirror go here)
set -euo pipefail
CONFIG_DIR=${1:-/kubeshare/scheduler/config}
PORT_DIR=${2:-/kubeshare/scheduler/podmanagerport}
LIB_DIR=${3:-/kubeshare/library}
HERE="$(cd "$(dirname "$0")" && pwd)"
ARBITER=${GEM_ARBITER:-$HERE/../bin/gem-arbiter}
BASE_QUOTA=${BASE_QUOTA:-300}; MIN_QUOTA=${MIN_QUOTA:-20}; WINDOW=${WINDOW:-10000}   # reference launcher.py:77-80

pids=()
trap 'kill "${pids[@]}" 2>/dev/null || true' EXIT INT TERM
i=0
for uuid in $(nvidia-smi --query-gpu=uuid --format=csv,noheader); do
  [ -f "$CONFIG_DIR/$uuid" ] || echo 0 > "$CONFIG_DIR/$uuid"     # reference launcher-multigpus.sh:26-31
  [ -f "$PORT_DIR/$uuid" ] || echo 0 > "$PORT_DIR/$uuid"
  "$ARBITER" --pool "$LIB_DIR/gemhook-$uuid.pool" -p "$CONFIG_DIR" -f "$uuid" --port-file "$PORT_DIR/$uuid" \
             --mirror "$LIB_DIR/config-$uuid" --columns limit_request -P $((49901 + i)) \
             -q "$BASE_QUOTA" -m "$MIN_QUOTA" -w "$WINDOW" &
  pids+=($!)
  i=$((i + 1))
done
wait
