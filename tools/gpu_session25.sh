#!/bin/bash
# stream-K remainder: targeted check under a short timeout first, then the GPU suite, bench A/B, vocoder A/B
set -u
O=gpurun_out
T=${1:-r2s25}
mkdir -p $O
export PYTHONUNBUFFERED=1
( timeout 240 python tools/tc_gemm_check.py streamk ) > $O/${T}_streamk.log 2>&1
rc=$?
echo "streamk rc=$rc" >> $O/${T}_streamk.log
if [ $rc -ne 0 ]; then
  echo "stream-K check failed: running the suite with SSB_STREAMK=0 only" >> $O/${T}_streamk.log
  export SSB_STREAMK=0
fi
( time timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 ) > $O/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> $O/${T}_pytest.log
( timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench.json 2> $O/${T}_bench.err
( SSB_STREAMK=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench_classic.json 2> $O/${T}_bench_classic.err
( timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench2.json 2> $O/${T}_bench2.err
( SSB_STREAMK=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-torch-leg --no-cpu --no-side ) > $O/${T}_bench_classic2.json 2> $O/${T}_bench_classic2.err
( timeout 600 python tools/vocoder_bench.py 600 --kernels --no-cpu ) > $O/${T}_vocoder_bench.json 2> $O/${T}_vocoder_bench.err
ls -la $O | grep ${T}
