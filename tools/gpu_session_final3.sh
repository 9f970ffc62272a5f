#!/bin/bash
# last verification of the round: GPU suite, smoke, default bench line, steady-state composition, cfg5
set -u
O=gpurun_out
T=${1:-r2h}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 300 python tools/step_profile.py --top 70 --steady 6 --seq $O/${T}_seq.tsv > $O/${T}_step_cupti.txt 2>&1
( time timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-torch-leg --no-cpu ) > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
ls -la $O | grep ${T}_
