#!/bin/bash
# GEMM epilogue (bias in smem, double-buffered tcgen05.ld), screener-only sample phase: tests, encode breakdown, launch lists, full bench
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_gemm test_gpu_towers test_gpu_train test_gpu_search test_gpu_hybrid test_gpu_service; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout=600 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $? $(tail -1 gpurun_out/$f.log)" >> gpurun_out/summary.txt
done
timeout 300 python tools/encode_bench.py 7680 7680 > gpurun_out/encode_bench.txt 2>&1
TTR_DEBUG_FLAGS=1073741824 timeout 300 python tools/encode_bench.py 7680 7680 > gpurun_out/encode_bench_single.txt 2>&1
TTR_DEBUG_FLAGS=524288 timeout 300 python tools/encode_bench.py 7680 7680 > gpurun_out/encode_bench_l0pair.txt 2>&1
for cfg in "8841823 256 0" "1105228 128 0" "1000000 256 0" "8841823 4096 0"; do set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --docs $1 --batch $2 --debug-flags $3 --no-extra --no-cpu-baseline > gpurun_out/sweep_d$1_b$2_f$3.log 2>&1
  echo "sweep $1 $2 $3 exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/sweep_d$1_b$2_f$3.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_b256.csv \
  python bench.py --steps 3 --warmup 3 --batch 256 --no-extra --no-cpu-baseline > gpurun_out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_d1105228_b128.csv \
  python bench.py --steps 3 --warmup 3 --docs 1105228 --batch 128 --no-extra --no-cpu-baseline > gpurun_out/ncu_l2.log 2>&1
timeout 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -4 gpurun_out/test_gpu_train.log | cut -c1-300
grep -h "input projection\|passages/s" gpurun_out/encode_bench.txt gpurun_out/encode_bench_single.txt gpurun_out/encode_bench_l0pair.txt
tail -3 gpurun_out/bench.err | cut -c1-300
