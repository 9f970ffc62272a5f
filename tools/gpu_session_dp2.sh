#!/bin/bash
# 2-GPU session: data-parallel bench + dp_check, cfg-5, reference arm
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/r2dp2_smi.txt 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-side ) > $O/r2dp2_bench_n2.json 2> $O/r2dp2_bench_n2.err
echo "rc=$?" >> $O/r2dp2_bench_n2.err
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-side --no-cpu --no-torch-leg ) > $O/r2dp2_bench_n1.json 2> $O/r2dp2_bench_n1.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 6 --warmup 3 --workload cfg5 ) > $O/r2dp2_bench_cfg5_n2.json 2> $O/r2dp2_bench_cfg5_n2.err
echo "rc=$?" >> $O/r2dp2_bench_cfg5_n2.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 ) > $O/r2dp2_bench_ref_n2.json 2> $O/r2dp2_bench_ref_n2.err
echo "rc=$?" >> $O/r2dp2_bench_ref_n2.err
ls -la $O | grep r2dp2
