#!/bin/bash
# tower-path pass: parity tests (each file in its own process, short timeouts), encode + train micro-benchmarks
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_gemm test_gpu_towers test_gpu_train test_gpu_service; do
  timeout 300 python -m pytest tests/$f.py -q -m gpu -x --timeout=120 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
timeout 120 python tools/trace_gru.py 3840 100 > gpurun_out/gru_trace.txt 2>&1
timeout 300 python tools/encode_bench.py 15360 7680 > gpurun_out/encode_bench.txt 2>&1; echo "encode_bench exit $?" >> gpurun_out/summary.txt
TTR_FP32_PIPELINE=1 timeout 300 python tools/encode_bench.py 15360 7680 > gpurun_out/encode_bench_fp32.txt 2>&1
timeout 300 python tools/train_bench.py > gpurun_out/train_bench.txt 2>&1; echo "train_bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -12 gpurun_out/test_gpu_gemm.log | cut -c1-200; tail -22 gpurun_out/test_gpu_towers.log | cut -c1-200; tail -3 gpurun_out/test_gpu_train.log gpurun_out/test_gpu_service.log | cut -c1-200
cat gpurun_out/encode_bench.txt gpurun_out/encode_bench_fp32.txt; head -4 gpurun_out/train_bench.txt
