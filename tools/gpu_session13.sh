#!/bin/bash
# Philox-7 check + evidence run: tests, full bench line, cfg-5, attention capture, EMG timing
set -u
O=gpurun_out
T=${1:-r2s13}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/${T}_bench.json 2> $O/${T}_bench.err
( time timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
timeout 300 python tools/step_profile.py --top 70 --seq $O/${T}_seq.tsv > $O/${T}_step_cupti.txt 2>&1
timeout 300 python tools/emg_bench.py 256 12000 > $O/${T}_emg_bench.json 2> $O/${T}_emg_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused -s 2 -c 2 -f -o $O/${T}_attn python tools/profile_targets.py attn 2 > $O/${T}_ncu_attn.log 2>&1
ls -la $O | grep ${T}
