#!/bin/bash
# 2-GPU box: deferred exchange (2-rank parity incl. the new stream-of-batches test), 2-GPU bench; on GPU 0: gemm tests (TMA store default)
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout=800 -s > gpurun_out/test_gpu_multi_n2.log 2>&1
echo "test_gpu_multi exit $? $(tail -1 gpurun_out/test_gpu_multi_n2.log)" >> gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_search.py -q -m gpu -x --timeout=800 > gpurun_out/test_gemm_search.log 2>&1
echo "gemm+search exit $? $(tail -1 gpurun_out/test_gemm_search.log)" >> gpurun_out/summary.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "bench N=2 exit $?" >> gpurun_out/summary.txt
for B in 1 256; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 3 --batch $B --no-extra > gpurun_out/bench_n2_b$B.log 2>&1
echo "bench N=2 B=$B exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/bench_n2_b$B.log | head -5 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; grep -E "world [0-9]+:" gpurun_out/test_gpu_multi_n2.log | cut -c1-400; tail -3 gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
for l in open('gpurun_out/bench_n2.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N=2 value ms', d['ms_per_step'], 'sync', d['synchronous']['ms_per_step'], 'local', d['roofline']['ms_per_call'], 'sust', d['sustained']['ms_per_step'], d['verified'])
PY
