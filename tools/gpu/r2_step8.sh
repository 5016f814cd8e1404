#!/bin/bash
# 2-GPU box: 2-rank parity, hybrid breakdown at 2 ranks, 2-GPU bench; on GPU 0: GEMM probe, sample-size sweep with the histogram bound
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout=800 -s > gpurun_out/test_gpu_multi_n2.log 2>&1
echo "test_gpu_multi exit $? $(tail -1 gpurun_out/test_gpu_multi_n2.log)" >> gpurun_out/summary.txt
for cfg in "8841823 4096" "8841823 256"; do set -- $cfg
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    tools/hybrid_breakdown_n.py $1 $2 > gpurun_out/hybrid_breakdown_n2_$1_$2.txt 2>&1
tail -1 gpurun_out/hybrid_breakdown_n2_$1_$2.txt >> gpurun_out/summary.txt
done
timeout 300 python tools/gemm_probe.py > gpurun_out/gemm_probe.txt 2>&1
cat gpurun_out/gemm_probe.txt >> gpurun_out/summary.txt
# sample tiles per SM: bits 12-17
for cfg in "1105228 128 0" "1105228 128 16384" "1105228 128 24576" "1105228 128 32768" "1105228 128 65536" "1000000 256 0" "1000000 256 16384" "1000000 256 24576" "1000000 256 65536" "8841823 128 0" "8841823 128 4194304"; do set -- $cfg
  timeout 300 python bench.py --steps 20 --warmup 3 --docs $1 --batch $2 --debug-flags $3 --no-extra --no-cpu-baseline > gpurun_out/sweep_d$1_b$2_f$3.log 2>&1
  echo "sweep $1 $2 $3 exit $? $(grep -h -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"verified": [a-z]*' gpurun_out/sweep_d$1_b$2_f$3.log | head -4 | tr '\n' ' ')" >> gpurun_out/summary.txt
done
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
echo "bench N=2 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt; tail -3 gpurun_out/bench_n2.err | cut -c1-300
