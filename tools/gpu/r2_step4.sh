#!/bin/bash
# screening epilogue: parity tests, batch sweep with and without screening (bit 29 = 536870912), encode leg, traces
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_search test_gpu_hybrid test_gpu_service; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout=600 > gpurun_out/$f.log 2>&1
  echo "$f exit $? $(tail -1 gpurun_out/$f.log)" >> gpurun_out/summary.txt
done
for cfg in "8841823 128 0" "8841823 128 536870912" "8841823 256 0" "8841823 256 536870912" "8841823 4096 0" "8841823 1024 0" "8841823 16 0" "1105228 128 0" "1105228 128 536870912" "1105228 4096 0" "1000000 256 0" "1105228 256 0"; do set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --docs $1 --batch $2 --debug-flags $3 --no-extra --no-cpu-baseline > gpurun_out/sweep_d$1_b$2_f$3.log 2>&1
  echo "sweep $1 $2 $3 exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/sweep_d$1_b$2_f$3.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
timeout 600 python bench.py --legs config3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1
echo "config3 $(grep -o '"passages_per_s": [0-9.]*\|"ms_wall": [0-9.]*\|"ms_device": [0-9.]*' gpurun_out/bench_c3.log | tr '\n' ' ')" >> gpurun_out/summary.txt
timeout 120 python tools/trace_scorer.py 256 0 4000000 > gpurun_out/trace5_b256_f0.txt 2>&1
cat gpurun_out/summary.txt; sed -n '1,3p;/^mean/,$p' gpurun_out/trace5_b256_f0.txt; tail -5 gpurun_out/test_gpu_search.log
timeout 300 python tools/encode_bench.py 7680 7680 > gpurun_out/plain_encode.log 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name regex:gemm_bias_kernel -c 4 -o gpurun_out/prof_r2_gemm -f python tools/encode_bench.py 7680 7680 > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out/prof_r2_gemm.ncu-rep
for f in test_gpu_gemm test_gpu_towers; do
  timeout 600 python -m pytest tests/$f.py -q -m gpu -x --timeout=300 > gpurun_out/$f.log 2>&1
  echo "$f exit $? $(tail -1 gpurun_out/$f.log)" | tee -a gpurun_out/summary.txt
done
timeout 300 python tools/encode_bench.py 7680 7680 > gpurun_out/plain_encode_pair.log 2>&1; tail -12 gpurun_out/plain_encode_pair.log
