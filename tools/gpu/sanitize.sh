#!/bin/bash
# compute-sanitizer over small shapes of the tcgen05 kernels (race / sync checking) -> profiles/r2_sanitizer_*.txt
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  for part in search towers train; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py $part > gpurun_out/sanitizer_${tool}_${part}.txt 2>&1
    echo "$tool $part exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_${part}.txt | tail -1)"
  done
done
