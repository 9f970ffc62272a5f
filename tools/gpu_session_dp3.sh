#!/bin/bash
# 2-GPU check of the final tree: data-parallel bench + dp_check
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/r2dp3_smi.txt 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-side ) > $O/r2dp3_bench_n2.json 2> $O/r2dp3_bench_n2.err
echo "rc=$?" >> $O/r2dp3_bench_n2.err
ls -la $O | grep r2dp3
