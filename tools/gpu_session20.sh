#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s20}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
( SSB_MASKBITS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench_nobits.json 2> $O/${T}_bench_nobits.err
( timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench.json 2> $O/${T}_bench.err
( SSB_MASKBITS=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench_nobits2.json 2> $O/${T}_bench_nobits2.err
( timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench2.json 2> $O/${T}_bench2.err
ls -la $O | grep ${T}
