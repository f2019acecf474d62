#!/bin/bash
# build the default library and the named variants; stop at the first failure
set -e
cd "$(dirname "$0")/.."
MLRN_QUIET=1 python -c "import __graft_entry__ as g; g.build()"
if [ $# -gt 0 ]; then python tests/build_variants.py "$@"; fi
