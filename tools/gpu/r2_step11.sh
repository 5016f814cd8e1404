#!/bin/bash
# 1 GPU: pipelined pair-GEMM epilogue (tests, probe incl. the un-pipelined A/B, encode breakdown)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_gemm test_gpu_towers; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout=600 > gpurun_out/$f.log 2>&1
  echo "$f exit $? $(tail -1 gpurun_out/$f.log)" >> gpurun_out/summary.txt
done
timeout 300 python tools/gemm_probe.py > gpurun_out/gemm_probe.txt 2>&1
grep "K=" gpurun_out/gemm_probe.txt >> gpurun_out/summary.txt
timeout 300 python tools/encode_bench.py 7680 7680 > gpurun_out/encode_bench.txt 2>&1
grep -h "input projection\|passages/s" gpurun_out/encode_bench.txt >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
