#!/bin/bash
# scorer timelines: CTA pairs vs one CTA per query tile
mkdir -p gpurun_out
for cfg in "256 0" "256 134217728" "1024 0" "1024 134217728" "256 33554432"; do set -- $cfg
  timeout 120 python tools/trace_scorer.py $1 $2 4000000 > gpurun_out/trace2_b$1_f$2.txt 2>&1
  echo "=== B=$1 flags=$2"; sed -n '1,3p;/^ 100/,$p' gpurun_out/trace2_b$1_f$2.txt
done
