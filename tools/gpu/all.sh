#!/bin/bash
# full GPU pass: every parity test file in its own process, smoke, default bench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_gemm test_gpu_search test_gpu_towers test_gpu_train test_gpu_hybrid test_gpu_service; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu -x --timeout=300 > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -4 gpurun_out/test_gpu_*.log | cut -c1-300
tail -3 gpurun_out/smoke.log; tail -2 gpurun_out/bench.log | cut -c1-3000; tail -1 gpurun_out/bench_ref.log | cut -c1-600
