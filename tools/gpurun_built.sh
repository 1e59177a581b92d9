#!/bin/bash
# Rebuild every native artefact, then hand the command to gpurun (built .so files travel).
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build()" > /tmp/build.log 2>&1 || { tail -30 /tmp/build.log; exit 1; }
exec gpurun "$@"
