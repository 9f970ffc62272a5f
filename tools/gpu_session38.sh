#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s38}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/${T}_bench.json 2> $O/${T}_bench.err
ls -la $O | grep ${T}
