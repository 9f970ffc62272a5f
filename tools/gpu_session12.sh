#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s12}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 300 python tools/dtw_bench.py 10000 5 > $O/${T}_dtw_bench.json 2> $O/${T}_dtw_bench.err
timeout 300 python tools/dtw_bench.py 16 20 > $O/${T}_dtw_bench16.json 2>> $O/${T}_dtw_bench.err
( SSB_CONVPAIR=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench_nopair.json 2> $O/${T}_bench_nopair.err
( timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 300 python tools/step_profile.py --top 70 --seq $O/${T}_seq.tsv > $O/${T}_step_cupti.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dtw_fill -s 1 -c 1 -f -o $O/${T}_dtw python tools/profile_targets.py dtw 2 > $O/${T}_ncu_dtw.log 2>&1
ls -la $O | grep ${T}
