#!/bin/bash
# ncu evidence (1 GPU): launch lists + one --set full capture of the dominant kernels (each command is first run without ncu)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 || exit 1
timeout 300 python tools/encode_bench.py 7680 7680 > gpurun_out/plain_encode.log 2>&1 || exit 1
timeout 300 python tools/train_bench.py > gpurun_out/plain_train.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_search.csv \
  python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_encode.csv \
  python tools/encode_bench.py 7680 7680 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_train.csv \
  python tools/train_bench.py > gpurun_out/ncu2b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-id ::regex:score_topk_mma_kernel:4 -c 1 \
  -o gpurun_out/prof_score_topk_mma -f python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-id ::regex:gru_fwd_tc_kernel:2 -c 1 \
  -o gpurun_out/prof_gru_fwd_tc -f python tools/encode_bench.py 7680 7680 > gpurun_out/ncu4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-id ::regex:gemm_bias_kernel:2 -c 1 \
  -o gpurun_out/prof_gemm_f16 -f python tools/encode_bench.py 7680 7680 > gpurun_out/ncu5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-id ::regex:gru_bwd_tc_kernel:3 -c 1 \
  -o gpurun_out/prof_gru_bwd_tc -f python tools/train_bench.py > gpurun_out/ncu6.log 2>&1
ls -la gpurun_out/*.ncu-rep
