#!/bin/bash
# N-GPU pass: multi-rank parity tests (2 / 4 / 8 ranks as available), then bench.py under torchrun at N GPUs
# usage: r2_multi.sh N [extra bench args]
N=${1:-2}; shift
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout=800 -s > gpurun_out/test_gpu_multi_n$N.log 2>&1
echo "test_gpu_multi exit $?" >> gpurun_out/summary.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 "$@" > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?" >> gpurun_out/summary.txt
for B in 1 256; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 3 --batch $B --no-extra > gpurun_out/bench_n${N}_b$B.log 2>&1
echo "bench N=$N B=$B exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/bench_n${N}_b$B.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
cat gpurun_out/summary.txt; grep -E "passed|failed|world [0-9]+:" gpurun_out/test_gpu_multi_n$N.log | cut -c1-1500; tail -5 gpurun_out/bench_n$N.err | cut -c1-400
tail -c 5000 gpurun_out/bench_n$N.log
