#!/bin/bash
# round-2 closing evidence session (after the vocoder / narrow-N MMA work): tests, smoke, bench lines,
# steady-state step composition, launch lists, ncu captures.  -> gpurun_out/r2g_*
set -u
O=gpurun_out
T=${1:-r2g}
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${T}_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/${T}_bench.json 2> $O/${T}_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
( time timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 --no-torch-leg ) > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
( timeout 600 python bench.py --eager --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench_eager.json 2> $O/${T}_bench_eager.err
timeout 300 python tools/step_profile.py --top 70 --steady 6 --seq $O/${T}_seq.tsv > $O/${T}_step_cupti.txt 2>&1
timeout 300 python tools/vocoder_bench.py 600 --kernels > $O/${T}_vocoder_bench.json 2> $O/${T}_vocoder_bench.err
# launch lists (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-torch-leg --no-cpu --no-side > $O/${T}_launches_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches_vocoder.csv python tools/vocoder_bench.py 600 --no-cpu > $O/${T}_launches_vocoder.log 2>&1
# full captures: the in-step FFN GEMMs (roofline.traffic), one narrow-N vocoder GEMM
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 4 -f -o $O/${T}_gemm python tools/profile_targets.py gemm 2 > $O/${T}_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 300 -c 3 -f -o $O/${T}_vocgemm python tools/vocoder_bench.py 600 --no-cpu > $O/${T}_ncu_vocgemm.log 2>&1
ls -la $O | grep ${T}_ | tail -40
