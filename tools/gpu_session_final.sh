#!/bin/bash
# round-2 evidence session: tests, bench lines, launch list, ncu captures, sanitizer.  -> gpurun_out/r2f_*
set -u
O=gpurun_out
T=${1:-r2f}
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${T}_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
( time timeout 900 python bench.py --steps 20 --warmup 5 ) > $O/${T}_bench.json 2> $O/${T}_bench.err
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
( time timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 ) > $O/${T}_bench_cfg5.json 2> $O/${T}_bench_cfg5.err
timeout 300 python tools/step_profile.py --top 70 --seq $O/${T}_seq.tsv > $O/${T}_step_cupti.txt 2>&1
timeout 300 python tools/emg_bench.py 256 12000 > $O/${T}_emg_bench.json 2> $O/${T}_emg_bench.err
timeout 300 python tools/dtw_bench.py 10000 5 > $O/${T}_dtw_bench.json 2> $O/${T}_dtw_bench.err
timeout 300 python tools/attn_bench.py fused > $O/${T}_attn_bench.txt 2>&1
# launch list of the bench command itself (cold-cache, serialised: shares, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${T}_launches.csv python bench.py --steps 2 --warmup 1 --no-torch-leg --no-cpu --no-side > $O/${T}_launches_bench.log 2>&1
for t in attn gemm; do
  case $t in attn) rx="attn_fused";; gemm) rx="gemm_tc_kernel";; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 4 -f -o $O/${T}_$t python tools/profile_targets.py $t 2 > $O/${T}_ncu_$t.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dtw_fill -s 1 -c 1 -f -o $O/${T}_dtw python tools/profile_targets.py dtw 2 > $O/${T}_ncu_dtw.log 2>&1
for tool in synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/attn_fused_check.py 1 200 2 96 99 > $O/${T}_san_${tool}_attn.log 2>&1
  echo "rc=$?" >> $O/${T}_san_${tool}_attn.log
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/tc_gemm_check.py ffn > $O/${T}_san_${tool}_gemm.log 2>&1
  echo "rc=$?" >> $O/${T}_san_${tool}_gemm.log
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/dtw_small_check.py > $O/${T}_san_${tool}_dtw.log 2>&1
  echo "rc=$?" >> $O/${T}_san_${tool}_dtw.log
done
ls -la $O | grep ${T}_ | tail -40
