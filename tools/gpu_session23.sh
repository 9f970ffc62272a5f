#!/bin/bash
# N-adaptive MMAs: full GPU suite, vocoder side bench (A/B), step composition with gap attribution, bench A/B
set -u
O=gpurun_out
T=${1:-r2s23}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
( timeout 600 python tools/vocoder_bench.py 600 --kernels ) > $O/${T}_vocoder_bench.json 2> $O/${T}_vocoder_bench.err
( SSB_TC_NARROW=0 timeout 600 python tools/vocoder_bench.py 600 --no-cpu ) > $O/${T}_vocoder_bench_wide.json 2> $O/${T}_vocoder_bench_wide.err
( timeout 600 python tools/step_profile.py --seq $O/${T}_seq.tsv ) > $O/${T}_step_cupti.txt 2>&1
( timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench.json 2> $O/${T}_bench.err
( SSB_TC_NARROW=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench_wide.json 2> $O/${T}_bench_wide.err
ls -la $O | grep ${T}
