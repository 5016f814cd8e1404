#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for B in 256 128; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 20 --warmup 3 --batch $B --no-extra > gpurun_out/bench_n2_b$B.log 2>&1
python - <<PY >> gpurun_out/summary.txt
import json
for l in open('gpurun_out/bench_n2_b$B.log'):
    if l.startswith('{'):
        d=json.loads(l); print('N=2 B=$B value ms', d['ms_per_step'], 'sync', d['synchronous']['ms_per_step'], 'local', d['roofline']['ms_per_call'], 'sust', d['sustained'], d['clocks'], d['verified'])
PY
done
cat gpurun_out/summary.txt
