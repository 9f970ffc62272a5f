#!/bin/bash
# vocoder bring-up: new GPU tests, the reference integration test (save_output with the drop-in Vocoder),
# the side bench with the per-kernel composition
set -u
O=gpurun_out
T=${1:-r2s22}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests/test_vocoder_gpu.py tests/test_reference_integration_gpu.py tests/test_ops_gpu.py -m gpu -q -s --maxfail=30 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
( timeout 600 python tools/vocoder_bench.py 600 --kernels ) > $O/${T}_vocoder_bench.json 2> $O/${T}_vocoder_bench.err
ls -la $O | grep ${T}
