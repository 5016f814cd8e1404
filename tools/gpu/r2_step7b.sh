#!/bin/bash
# 1 GPU: all parity tests (weight-stationary pair GEMM, survivor histogram, 8 accumulators, lanes in encode_rows / train_step),
# encode + train breakdowns, small-shard traces, search sweeps with the histogram bound on / off
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_gemm test_gpu_search test_gpu_towers test_gpu_train test_gpu_hybrid test_gpu_service; do
  timeout 1200 python -m pytest tests/$f.py -q -m gpu -x --timeout=900 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $? $(tail -1 gpurun_out/$f.log)" >> gpurun_out/summary.txt
done
timeout 300 python tools/trace_scorer.py 128 0 1105228 all > gpurun_out/trace7_b128_1p1M.txt 2>&1
timeout 300 python tools/trace_scorer.py 256 0 1000000 all > gpurun_out/trace7_b256_1M.txt 2>&1
for cfg in "1105228 128 0" "1105228 128 -2147483648" "1000000 256 0" "1000000 256 -2147483648" "8841823 128 0" "8841823 256 0" "8841823 256 -2147483648" "1105228 1 0" "1105228 512 0"; do set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 3 --docs $1 --batch $2 --debug-flags $3 --no-extra --no-cpu-baseline > gpurun_out/sweep_d$1_b$2_f$3.log 2>&1
  echo "sweep $1 $2 $3 exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/sweep_d$1_b$2_f$3.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
