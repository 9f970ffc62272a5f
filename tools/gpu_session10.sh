#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s10}
mkdir -p $O
export PYTHONUNBUFFERED=1
( time timeout 900 python -m pytest tests/test_emg_gpu.py tests/test_dtw_gpu.py tests/test_fused_loss_gpu.py -q --maxfail=10 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 300 python tools/dtw_bench.py 10000 5 > $O/${T}_dtw_bench.json 2> $O/${T}_dtw_bench.err
timeout 300 python tools/dtw_bench.py 16 20 > $O/${T}_dtw_bench16.json 2>> $O/${T}_dtw_bench.err
timeout 300 python tools/emg_bench.py 256 12000 > $O/${T}_emg_bench.json 2> $O/${T}_emg_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dtw_fill -s 1 -c 1 -f -o $O/${T}_dtw python tools/profile_targets.py dtw 2 > $O/${T}_ncu_dtw.log 2>&1
ls -la $O | grep ${T}
