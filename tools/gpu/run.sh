#!/bin/bash
# usage: tools/gpu/run.sh <test files...>   (each file in its own process)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in "$@"; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu --timeout=600 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
