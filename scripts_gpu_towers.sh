#!/bin/bash
# tower-path pass: parity tests (each file in its own process, short timeouts), encode + train micro-benchmarks
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_towers test_gpu_train; do
  timeout 300 python -m pytest tests/$f.py -q -m gpu -x --timeout=120 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
timeout 120 python tools/trace_gru.py 4096 100 > gpurun_out/gru_trace.txt 2>&1
timeout 300 python tools/encode_bench.py > gpurun_out/encode_bench.txt 2>&1; echo "encode_bench exit $?" >> gpurun_out/summary.txt
timeout 300 python tools/train_bench.py > gpurun_out/train_bench.txt 2>&1; echo "train_bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -15 gpurun_out/test_gpu_towers.log; tail -5 gpurun_out/test_gpu_train.log
cat gpurun_out/gru_trace.txt | tail -22
cat gpurun_out/encode_bench.txt; head -8 gpurun_out/train_bench.txt
