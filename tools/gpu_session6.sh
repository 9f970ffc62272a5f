#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 ) > $O/r2s6_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2s6_pytest.log
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-torch-leg --no-cpu --no-side ) > $O/r2s6_bench.json 2> $O/r2s6_bench.err
timeout 300 python tools/step_profile.py --top 70 --seq $O/r2s6_seq.tsv > $O/r2s6_step_cupti.txt 2>&1
ls -la $O | grep r2s6
