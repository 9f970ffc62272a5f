#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s9}
mkdir -p $O
export PYTHONUNBUFFERED=1
for f in 0 1 2 3; do
  SSB_DTW_FLAGS=$f timeout 300 python tools/dtw_bench.py 10000 5 > $O/${T}_dtw_f$f.json 2>> $O/${T}_dtw.err
done
( SSB_DTW_FLAGS=3 timeout 900 python -m pytest tests/test_dtw_gpu.py -q --maxfail=10 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
SSB_DTW_FLAGS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dtw_fill -s 1 -c 1 -f -o $O/${T}_dtw python tools/profile_targets.py dtw 2 > $O/${T}_ncu_dtw.log 2>&1
ls -la $O | grep ${T}
