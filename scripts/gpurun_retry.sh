#!/bin/bash
# usage: scripts/gpurun_retry.sh <gpus> <timeout_s> <command...>   — retries while the pod answers busy (rc 3)
G=$1; T=$2; shift 2
for i in $(seq 1 40); do
  if [ "$G" = "1" ]; then /usr/local/graft/bin/gpurun --timeout $T -- "$@"; else /usr/local/graft/bin/gpurun --gpus $G --timeout $T -- "$@"; fi
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[retry] busy, attempt $i"; sleep 45
done
exit 3
