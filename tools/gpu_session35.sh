#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s37}
mkdir -p $O
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_parity_bench_engine_gpu.py tests/test_graph_gpu.py -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 300 python tools/step_profile.py --top 70 --steady 6 > $O/${T}_step_cupti.txt 2>&1
ls -la $O | grep ${T}
