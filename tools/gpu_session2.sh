#!/bin/bash
# round-2 GPU session 2: new kernels (ragged DTW, fused loss, CTC v2, fp64 DTW), integration test
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests/test_fused_loss_gpu.py tests/test_ctc_gpu.py tests/test_dtw_gpu.py tests/test_load_audio_gpu.py tests/test_reference_integration_gpu.py tests/test_step_gpu.py tests/test_graph_gpu.py tests/test_recognition_gpu.py -q -s --maxfail=30 ) > $O/r2s2_pytest_new.log 2>&1
echo "pytest rc=$?" >> $O/r2s2_pytest_new.log
( time timeout 1200 python -m pytest tests -m gpu -q --maxfail=25 --deselect tests/test_fused_loss_gpu.py --deselect tests/test_ctc_gpu.py --deselect tests/test_dtw_gpu.py --deselect tests/test_load_audio_gpu.py --deselect tests/test_reference_integration_gpu.py --deselect tests/test_step_gpu.py --deselect tests/test_graph_gpu.py --deselect tests/test_recognition_gpu.py ) > $O/r2s2_pytest_rest.log 2>&1
echo "pytest rc=$?" >> $O/r2s2_pytest_rest.log
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-torch-leg ) > $O/r2s2_bench.json 2> $O/r2s2_bench.err
( time timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-torch-leg --no-cpu ) > $O/r2s2_bench_cfg5.json 2> $O/r2s2_bench_cfg5.err
timeout 300 python tools/step_profile.py --top 70 > $O/r2s2_step_cupti.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mel_kernel -s 1 -c 1 -f -o $O/r2s2_mel python tools/profile_targets.py mel 2 > $O/r2s2_ncu_mel.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:ctc_ -c 3 -f -o $O/r2s2_ctc python tools/profile_targets.py ctc 1 > $O/r2s2_ncu_ctc.log 2>&1
ls -la $O | grep r2s2
