#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_hybrid.py -q -m gpu -x --timeout=800 > gpurun_out/test_search.log 2>&1
echo "search+hybrid exit $? $(tail -1 gpurun_out/test_search.log)" >> gpurun_out/summary.txt
for cfg in "1000000 1 0" "1105228 1 0" "8841823 1 0" "1000000 4 0" "1000000 2 0"; do set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 3 --docs $1 --batch $2 --debug-flags $3 --no-extra --no-cpu-baseline > gpurun_out/sweep_d$1_b$2_f$3.log 2>&1
  echo "sweep $1 $2 $3 exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/sweep_d$1_b$2_f$3.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
