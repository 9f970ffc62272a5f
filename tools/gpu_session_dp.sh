#!/bin/bash
# 2-GPU session: the data-parallel path (overlapped segmented all-reduce captured in the graph + dp_check)
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi -L > $O/r2dp_smi.txt 2>&1
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-side ) > $O/r2dp_bench_n2.json 2> $O/r2dp_bench_n2.err
echo "rc=$?" >> $O/r2dp_bench_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-side --no-overlap ) > $O/r2dp_bench_n2_noov.json 2> $O/r2dp_bench_n2_noov.err
echo "rc=$?" >> $O/r2dp_bench_n2_noov.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 6 --warmup 3 --workload cfg5 ) > $O/r2dp_bench_cfg5_n2.json 2> $O/r2dp_bench_cfg5_n2.err
echo "rc=$?" >> $O/r2dp_bench_cfg5_n2.err
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 ) > $O/r2dp_bench_ref_n2.json 2> $O/r2dp_bench_ref_n2.err
echo "rc=$?" >> $O/r2dp_bench_ref_n2.err
ls -la $O | grep r2dp
