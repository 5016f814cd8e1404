#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for cfg in "1105228 0" "1105228 100000" "1600000 0" "1600000 100000" "1105228 0" "1105228 100000" "800000 0" "800000 100000"; do set -- $cfg
  TTR_HIST_MAX_TILES=$2 timeout 300 python bench.py --steps 20 --warmup 3 --docs $1 --batch 128 --no-extra --no-cpu-baseline > gpurun_out/hist_d$1_t$2.log 2>&1
  echo "docs $1 hist_max_tiles $2 exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/hist_d$1_t$2.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt
