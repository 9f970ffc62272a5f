#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s29}
mkdir -p $O
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_gemm_narrow_gpu.py -m gpu -q --maxfail=60 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
ls -la $O | grep ${T}
