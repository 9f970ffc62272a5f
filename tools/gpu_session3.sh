#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests/test_ctc_gpu.py tests/test_fused_loss_gpu.py tests/test_reference_integration_gpu.py tests/test_parity_bench_engine_gpu.py tests/test_graph_gpu.py tests/test_recognition_gpu.py -q --maxfail=30 ) > $O/r2s3_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2s3_pytest.log
( time timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-torch-leg --no-cpu ) > $O/r2s3_bench_cfg5.json 2> $O/r2s3_bench_cfg5.err
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-torch-leg --no-cpu --no-side ) > $O/r2s3_bench.json 2> $O/r2s3_bench.err
timeout 300 python tools/step_profile.py --top 70 --seq $O/r2s3_seq.tsv > $O/r2s3_step_cupti.txt 2>&1
timeout 300 ncu --set full --clock-control none -k regex:ctc_ -c 3 -f -o $O/r2s3_ctc python tools/profile_targets.py ctc 1 > $O/r2s3_ncu_ctc.log 2>&1
ls -la $O | grep r2s3
