#!/bin/bash
# round-2 GPU session 1: tests, benches, captures.  Everything lands in gpurun_out/.
set -u
O=gpurun_out
mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2s1_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 -x -k "not reference_integration and not parity_bench" ) > $O/r2s1_pytest.log 2>&1
echo "pytest rc=$?" >> $O/r2s1_pytest.log
( time timeout 900 python -m pytest tests/test_reference_integration_gpu.py tests/test_parity_bench_engine_gpu.py -q -s ) > $O/r2s1_pytest_new.log 2>&1
echo "pytest rc=$?" >> $O/r2s1_pytest_new.log
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > $O/r2s1_bench.json 2> $O/r2s1_bench.err
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r2s1_bench_ref.json 2> $O/r2s1_bench_ref.err
( time timeout 600 python bench.py --workload cfg5 --steps 10 --warmup 3 ) > $O/r2s1_bench_cfg5.json 2> $O/r2s1_bench_cfg5.err
timeout 300 python tools/step_profile.py --top 60 > $O/r2s1_step_cupti.txt 2>&1
# ncu: full captures of the kernels VERDICT asks about
for t in attn mel gemm; do
  case $t in attn) rx="attn_fused";; mel) rx="mel_kernel";; gemm) rx="gemm_tc_kernel";; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 4 -f -o $O/r2s1_$t python tools/profile_targets.py $t 2 > $O/r2s1_ncu_$t.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dtw_fill -s 1 -c 1 -f -o $O/r2s1_dtw python tools/profile_targets.py dtw 2 > $O/r2s1_ncu_dtw.log 2>&1
# sanitizer: synccheck + racecheck over the hand-rolled mbarrier / TMA / TMEM pipelines
for tool in synccheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/attn_fused_check.py 1 200 2 96 99 > $O/r2s1_san_${tool}_attn.log 2>&1
  echo "rc=$?" >> $O/r2s1_san_${tool}_attn.log
  timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/tc_gemm_check.py ffn > $O/r2s1_san_${tool}_gemm.log 2>&1
  echo "rc=$?" >> $O/r2s1_san_${tool}_gemm.log
done
ls -la $O | tail -40
