#!/bin/bash
# search-path pass: parity tests, bench at full and 1/8 corpus (fused and two-pass), ncu launch list
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_search.py tests/test_gpu_hybrid.py -q -m gpu -x --timeout=300 > gpurun_out/test_gpu_search.log 2>&1
echo "test_gpu_search exit $?" >> gpurun_out/summary.txt
for cfg in "8841823 128 0" "8841823 128 2097152" "8841823 16 0" "1105228 128 0" "1105228 128 2097152" "1105228 16 0"; do
  set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --docs $1 --batch $2 --debug-flags $3 --no-extra --no-cpu-baseline > gpurun_out/bench_d$1_b$2_f$3.log 2>&1
  echo "bench $1 $2 $3 exit $?" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; tail -3 gpurun_out/test_gpu_search.log
for f in gpurun_out/bench_d*_b*_f*.log; do echo $f; grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"ms_per_call": [0-9.]*' $f | tr '\n' ' '; echo; done
