#!/bin/bash
# query-batch sweep on one GPU (full corpus), for profiles/
mkdir -p gpurun_out
for B in 1 4 8 16 64 128 256 1024 4096; do
  timeout 300 python bench.py --steps 10 --warmup 3 --batch $B --no-extra --no-cpu-baseline > gpurun_out/sweep_b$B.log 2>&1
  echo "B=$B $(grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"achieved": [0-9.]*\|"frac": [0-9.]*' gpurun_out/sweep_b$B.log | tr '\n' ' ')"
done
