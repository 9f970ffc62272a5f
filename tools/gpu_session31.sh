#!/bin/bash
# packed FFMA2 in the thin-K kernels: parity tests + per-kernel times in the step
set -u
O=gpurun_out
T=${1:-r2s33}
mkdir -p $O
export PYTHONUNBUFFERED=1
( timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_model_gpu.py tests/test_parity_bench_engine_gpu.py -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 300 python tools/step_profile.py --top 70 --steady 6 > $O/${T}_step_cupti.txt 2>&1
( timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench.json 2> $O/${T}_bench.err
ls -la $O | grep ${T}
