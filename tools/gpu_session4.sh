#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests/test_weight_planes_gpu.py tests/test_model_gpu.py tests/test_graph_gpu.py tests/test_step_gpu.py tests/test_parity_bench_engine_gpu.py tests/test_recognition_gpu.py tests/test_optim_gpu.py -q --maxfail=30 ) > $O/r2s4_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2s4_pytest.log
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-torch-leg --no-cpu --no-side ) > $O/r2s4_bench.json 2> $O/r2s4_bench.err
SSB_WPLANES=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-torch-leg --no-cpu --no-side > $O/r2s4_bench_nowp.json 2> $O/r2s4_bench_nowp.err
timeout 300 python tools/step_profile.py --top 70 --seq $O/r2s4_seq.tsv > $O/r2s4_step_cupti.txt 2>&1
ls -la $O | grep r2s4
