#!/bin/bash
# vocoder: side streams on every stage (tail waves of one branch under the next branch's kernels) vs underfilled stages only
set -u
O=gpurun_out
T=${1:-r2s27}
mkdir -p $O
export PYTHONUNBUFFERED=1
for m in 2 1 2 1; do
( SSB_VOC_STREAMS=$m timeout 600 python tools/vocoder_bench.py 600 --no-cpu ) >> $O/${T}_vocoder_bench_m$m.json 2>> $O/${T}_vocoder_bench_m$m.err
done
( SSB_VOC_STREAMS=2 timeout 600 python tools/vocoder_bench.py 611 --no-cpu ) > $O/${T}_vocoder_bench611_m2.json 2> $O/${T}_vocoder_bench611_m2.err
( SSB_VOC_STREAMS=2 time timeout 900 python -m pytest tests/test_vocoder_gpu.py -m gpu -q --maxfail=30 ) > $O/${T}_pytest.log 2>&1
ls -la $O | grep ${T}
