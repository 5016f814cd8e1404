#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:score_topk_mma_kernel -c 2 \
  -o gpurun_out/prof_r2_pair_b256 -f python bench.py --steps 1 --warmup 0 --batch 256 --no-extra --no-cpu-baseline > gpurun_out/ncu_pair.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:score_topk_mma_kernel -c 2 \
  -o gpurun_out/prof_r2_single_b256 -f python bench.py --steps 1 --warmup 0 --batch 256 --debug-flags 134217728 --no-extra --no-cpu-baseline > gpurun_out/ncu_single.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/ncu_pair.log
