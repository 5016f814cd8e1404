#!/bin/bash
# compute-sanitizer over small shapes of the tcgen05 kernels (race / sync checking) -> gpurun_out/sanitizer_*.txt
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for tool in memcheck synccheck racecheck; do
  for part in search towers train; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_driver.py $part > gpurun_out/sanitizer_${tool}_${part}.txt 2>&1
    echo "$tool $part exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitizer_${tool}_${part}.txt | tail -1) | $(grep -E '^(search|encode|train) ' gpurun_out/sanitizer_${tool}_${part}.txt | tr '\n' ';')" >> gpurun_out/summary.txt
  done
done
cat gpurun_out/summary.txt
