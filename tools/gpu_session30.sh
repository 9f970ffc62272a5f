#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s30}
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:thin_ -s 4 -c 4 -f -o $O/${T}_thin python tools/profile_targets.py thin 2 > $O/${T}_ncu_thin.log 2>&1
ls -la $O | grep ${T}
