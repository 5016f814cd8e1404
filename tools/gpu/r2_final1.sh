#!/bin/bash
# 1 GPU: the whole GPU suite, the default bench line with every leg, the reference arm, encode batch-size sweep
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -x --timeout=900 > gpurun_out/test_gpu_all.log 2>&1
echo "pytest -m gpu exit $? $(tail -1 gpurun_out/test_gpu_all.log)" >> gpurun_out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
echo "smoke exit $? $(tail -1 gpurun_out/smoke.log)" >> gpurun_out/summary.txt
timeout 1500 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?" >> gpurun_out/summary.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "reference arm exit $?" >> gpurun_out/summary.txt
timeout 600 python tools/encode_rows_profile.py 1105228 > gpurun_out/encode_rows_profile.txt 2>&1
grep lanes gpurun_out/encode_rows_profile.txt >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/bench.err | cut -c1-300; cat gpurun_out/bench_reference.log | cut -c1-600
