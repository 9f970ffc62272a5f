#!/bin/bash
# vocoder: residual branches on side streams where the grid is underfilled (A/B), tests
set -u
O=gpurun_out
T=${1:-r2s26}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests/test_vocoder_gpu.py tests/test_reference_integration_gpu.py -m gpu -q --maxfail=30 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
( timeout 600 python tools/vocoder_bench.py 600 --kernels --no-cpu ) > $O/${T}_vocoder_bench.json 2> $O/${T}_vocoder_bench.err
( SSB_VOC_STREAMS=0 timeout 600 python tools/vocoder_bench.py 600 --no-cpu ) > $O/${T}_vocoder_bench_1stream.json 2> $O/${T}_vocoder_bench_1stream.err
( timeout 600 python tools/vocoder_bench.py 150 --no-cpu ) > $O/${T}_vocoder_bench150.json 2> $O/${T}_vocoder_bench150.err
( SSB_VOC_STREAMS=0 timeout 600 python tools/vocoder_bench.py 150 --no-cpu ) > $O/${T}_vocoder_bench150_1stream.json 2> $O/${T}_vocoder_bench150_1stream.err
ls -la $O | grep ${T}
