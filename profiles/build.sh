#!/bin/bash
# build libnbody_b200.so + the oracle from any working directory; fails loudly
set -e
cd "$(dirname "$0")/.."
python -c "import __graft_entry__ as g; g.build(); print('build ok')"
