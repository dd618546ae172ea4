#!/bin/bash
# build + sanity-check locally, then run a command on the GPU box:  tools/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python -m fresco_b200.build > /tmp/fresco_build.log 2>&1 || { tail -20 /tmp/fresco_build.log; exit 1; }
test -f fresco_b200/libfresco_b200.so || { echo "libfresco_b200.so missing"; exit 1; }
python -c "import ctypes; assert ctypes.CDLL('fresco_b200/libfresco_b200.so').fresco_abi_version() == 1"
exec /usr/local/graft/bin/gpurun --timeout "$1" -- "$2"
