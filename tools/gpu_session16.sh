#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s16}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 300 python tools/attn_bench.py fused > $O/${T}_attn_bench.txt 2>&1
( timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench.json 2> $O/${T}_bench.err
ls -la $O | grep ${T}
