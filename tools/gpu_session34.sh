#!/bin/bash
set -u
O=gpurun_out
T=${1:-r2s34}
mkdir -p $O
export PYTHONUNBUFFERED=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:add_ln_ -s 2 -c 2 -f -o $O/${T}_ln python tools/profile_targets.py ln 2 > $O/${T}_ncu_ln.log 2>&1
ls -la $O | grep ${T}
