#!/bin/bash
# verification of the final tree: GPU suite, smoke, default bench line
set -u
O=gpurun_out
T=${1:-r2s28}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/${T}_bench.json 2> $O/${T}_bench.err
ls -la $O | grep ${T}
