#!/bin/bash
# 1 GPU: frontend-fixture service test, lanes in encode_rows, weight-stationary pair GEMM, encode profile, scorer traces on small shards
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_gemm test_gpu_service test_gpu_towers; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout=600 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $? $(tail -1 gpurun_out/$f.log)" >> gpurun_out/summary.txt
done
timeout 300 python tools/encode_bench.py 7680 7680 > gpurun_out/encode_bench.txt 2>&1
TTR_DEBUG_FLAGS=524288 timeout 300 python tools/encode_bench.py 7680 7680 > gpurun_out/encode_bench_streamw.txt 2>&1
grep -h "input projection\|passages/s" gpurun_out/encode_bench.txt gpurun_out/encode_bench_streamw.txt >> gpurun_out/summary.txt
timeout 600 python tools/encode_rows_profile.py 400000 > gpurun_out/encode_rows_profile.txt 2>&1
grep lanes gpurun_out/encode_rows_profile.txt >> gpurun_out/summary.txt
timeout 300 python tools/trace_scorer.py 128 0 1105228 all > gpurun_out/trace6_b128_1p1M.txt 2>&1
timeout 300 python tools/trace_scorer.py 256 0 1000000 all > gpurun_out/trace6_b256_1M.txt 2>&1
cat gpurun_out/summary.txt
