#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s14}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py -q --maxfail=10 -k "attention or attn or model" ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 300 python tools/attn_bench.py fused > $O/${T}_attn_bench.txt 2>&1
ls -la $O | grep ${T}
