#!/bin/bash
# CTA-pair scorer bring-up: diagnostic, search/hybrid/service tests, batch sweep
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python tools/check_pair.py > gpurun_out/check_pair.txt 2>&1; echo "check_pair exit $?" >> gpurun_out/summary.txt
for f in test_gpu_search test_gpu_hybrid test_gpu_service; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout=600 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
for cfg in "8841823 256 0" "8841823 256 134217728" "8841823 256 268435456" "8841823 4096 0" "8841823 1024 0" "1105228 4096 0" "1000000 256 0" "1000000 256 2097152" "1105228 256 0"; do set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --docs $1 --batch $2 --debug-flags $3 --no-extra --no-cpu-baseline > gpurun_out/sweep_d$1_b$2_f$3.log 2>&1
  echo "sweep $1 $2 $3 exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/sweep_d$1_b$2_f$3.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; cat gpurun_out/check_pair.txt
tail -n 5 gpurun_out/test_gpu_search.log gpurun_out/test_gpu_hybrid.log gpurun_out/test_gpu_service.log | cut -c1-300
timeout 120 python tools/trace_scorer.py 256 0 4000000 > gpurun_out/trace3_b256_f0.txt 2>&1; sed -n '1,3p;/^ 100/,$p' gpurun_out/trace3_b256_f0.txt
