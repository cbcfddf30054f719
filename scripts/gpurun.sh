#!/bin/bash
# builder convenience: rebuild the library (stale .so files travel as they are), then run a command on the GPU box
set -e
cd "$(dirname "$0")/.."
python buddy_b200/build.py > /dev/null
exec /usr/local/graft/bin/gpurun "$@"
