#!/bin/bash
# round-2 pass: every parity test file in its own process, smoke, scorer batch sweep, full default bench (all legs)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_search test_gpu_hybrid test_gpu_gemm test_gpu_towers test_gpu_train test_gpu_service; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout=600 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary.txt
for cfg in "8841823 128" "8841823 256" "8841823 4096" "1105228 128" "1000000 256" "1000000 1"; do set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --docs $1 --batch $2 --no-extra --no-cpu-baseline > gpurun_out/sweep_d$1_b$2.log 2>&1
  echo "sweep $1 $2 exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/sweep_d$1_b$2.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
timeout 1200 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/test_gpu_*.log | cut -c1-400
tail -3 gpurun_out/smoke.log; tail -5 gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.log
