#!/bin/bash
# Always rebuild in-tree binaries before shipping the snapshot to the GPU box.
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -20 /tmp/build.log; exit 1; }
exec /usr/local/graft/bin/gpurun "$@"
