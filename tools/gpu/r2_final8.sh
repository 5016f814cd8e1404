#!/bin/bash
# 8-GPU box, final code: multi-rank parity (2 / 4 / 8 ranks) and the default bench line at 8 GPUs
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout=500 -s > gpurun_out/test_gpu_multi_n8.log 2>&1
echo "test_gpu_multi exit $? $(tail -1 gpurun_out/test_gpu_multi_n8.log)" >> gpurun_out/summary.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err
echo "bench N=8 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -2 gpurun_out/bench_n8.err | cut -c1-300
