#!/bin/bash
# N-GPU pass: 2-rank parity tests, then the bench at the box's GPU count
# usage: multi.sh N [quick]   (quick: only the default bench line)
N=${1:-2}
QUICK=${2:-}
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
if [ -z "$QUICK" ]; then
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu -x --timeout=500 > gpurun_out/test_gpu_multi.log 2>&1
echo "test_gpu_multi exit $?" >> gpurun_out/summary.txt
fi
for B in $( [ -z "$QUICK" ] && echo 128 1 4096 ); do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 --batch $B --no-extra > gpurun_out/bench${N}_b$B.log 2>&1
  echo "bench N=$N B=$B exit $?" >> gpurun_out/summary.txt
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench${N}_default.log 2>&1
echo "bench N=$N default exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/test_gpu_multi.log
for f in gpurun_out/bench${N}_*.log; do echo $f; grep -h -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"passages_per_s": [0-9.]*' $f | tr '\n' ' '; echo; done
