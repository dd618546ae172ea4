#!/bin/bash
# tools/gpu_retry.sh <timeout_s> '<command>' : like tools/gpu.sh, retrying while the pod answers "no box / slot free" (rc 3)
cd "$(dirname "$0")/.."
for i in $(seq 1 40); do
  tools/gpu.sh "$1" "$2"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
