#!/bin/bash
# 1 GPU: same-box A/B of the histogram bound on the large shapes, encode_rows after the host-prep fix, ncu of the K = 200 pair GEMM
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for cfg in "4420912 4096 0" "4420912 4096 -2147483648" "8841823 128 0" "8841823 128 -2147483648" "8841823 128 0" "8841823 4096 0" "8841823 4096 -2147483648" "8841823 1024 0" "8841823 1024 -2147483648"; do set -- $cfg
  timeout 300 python bench.py --steps 10 --warmup 3 --docs $1 --batch $2 --debug-flags $3 --no-extra --no-cpu-baseline > gpurun_out/ab_d$1_b$2_f$3.log 2>&1
  echo "ab $1 $2 $3 exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*\|"sm_mhz": [0-9.]*' gpurun_out/ab_d$1_b$2_f$3.log | head -5 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
timeout 300 python -m pytest tests/test_gpu_towers.py tests/test_gpu_service.py -q -m gpu -x --timeout=600 > gpurun_out/test_enc.log 2>&1
echo "towers+service exit $? $(tail -1 gpurun_out/test_enc.log)" >> gpurun_out/summary.txt
timeout 600 python tools/encode_rows_profile.py 600000 > gpurun_out/encode_rows_profile.txt 2>&1
grep lanes gpurun_out/encode_rows_profile.txt >> gpurun_out/summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bias_pair -s 6 -c 2 -o gpurun_out/r2_gemm_pair_k200 python tools/gemm_probe.py 504769 > gpurun_out/ncu_gemm.log 2>&1
echo "ncu gemm exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
