#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu -x --timeout=120 > gpurun_out/test_gpu_gemm.log 2>&1
echo "test_gpu_gemm exit $? $(tail -1 gpurun_out/test_gpu_gemm.log)" >> gpurun_out/summary.txt
timeout 120 python tools/gemm_probe.py > gpurun_out/gemm_probe.txt 2>&1
grep "K=" gpurun_out/gemm_probe.txt >> gpurun_out/summary.txt
timeout 200 python -m pytest tests/test_gpu_towers.py -q -m gpu -x --timeout=120 > gpurun_out/test_gpu_towers.log 2>&1
echo "test_gpu_towers exit $? $(tail -1 gpurun_out/test_gpu_towers.log)" >> gpurun_out/summary.txt
timeout 200 python tools/encode_bench.py 7680 7680 > gpurun_out/encode_bench.txt 2>&1
grep -h "input projection\|passages/s" gpurun_out/encode_bench.txt >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -5 gpurun_out/test_gpu_gemm.log | cut -c1-300
