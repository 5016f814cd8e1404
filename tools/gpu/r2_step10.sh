#!/bin/bash
# 8-GPU box: multi-rank parity (2 / 4 / 8 ranks), bench at 8 and 4 GPUs
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
nvidia-smi -L | wc -l >> gpurun_out/summary.txt
timeout 1200 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout=1000 -s > gpurun_out/test_gpu_multi_n8.log 2>&1
echo "test_gpu_multi exit $? $(tail -1 gpurun_out/test_gpu_multi_n8.log)" >> gpurun_out/summary.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8.log 2> gpurun_out/bench_n8.err
echo "bench N=8 exit $?" >> gpurun_out/summary.txt
for B in 1 256; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 8 --steps 20 --warmup 3 --batch $B --no-extra > gpurun_out/bench_n8_b$B.log 2>&1
echo "bench N=8 B=$B exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/bench_n8_b$B.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 \
    bench.py --gpus 4 --steps 20 --warmup 3 --no-extra > gpurun_out/bench_n4.log 2> gpurun_out/bench_n4.err
echo "bench N=4 exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/bench_n4.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; grep -E "passed|failed|world [0-9]+:" gpurun_out/test_gpu_multi_n8.log | cut -c1-600; tail -3 gpurun_out/bench_n8.err | cut -c1-300
