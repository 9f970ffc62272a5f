#!/bin/bash
# usage: tools/gpurun_retry.sh <label> [--gpus N] <timeout> <command...>   (retries while the pod answers busy)
L=$1; shift
G=""
if [ "$1" = "--gpus" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun $G --timeout $T -- "$@" > gpurun_out/${L}_call.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
