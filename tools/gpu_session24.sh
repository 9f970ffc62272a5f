#!/bin/bash
# warp-per-run conv_post kernel, vocoder kernel sequence, steady-state step composition
set -u
O=gpurun_out
T=${1:-r2s24}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests/test_vocoder_gpu.py tests/test_reference_integration_gpu.py -m gpu -q -s --maxfail=30 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
( timeout 600 python tools/vocoder_bench.py 600 --kernels --no-cpu ) > $O/${T}_vocoder_bench.json 2> $O/${T}_vocoder_bench.err
( timeout 600 python tools/vocoder_bench.py 611 --no-cpu ) > $O/${T}_vocoder_bench611.json 2> $O/${T}_vocoder_bench611.err
( timeout 600 python tools/step_profile.py --steady 6 --seq $O/${T}_seq.tsv ) > $O/${T}_step_cupti_steady.txt 2>&1
ls -la $O | grep ${T}
