#!/bin/bash
# scorer pipeline diagnostics: per-tile timelines of CTA (0,0) at several query batches / debug flags, and the
# DRAM traffic + L2 hit rate of the main pass when several query tiles share document tiles through L2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for cfg in "128 0" "128 64" "128 8" "256 0" "1024 0" "1024 8"; do set -- $cfg
  timeout 120 python tools/trace_scorer.py $1 $2 4000000 > gpurun_out/trace_b$1_f$2.txt 2>&1
done
for B in 128 256 1024; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none --kernel-name regex:score_topk_mma_kernel -c 6 --csv --log-file gpurun_out/ncu_scorer_b$B.csv \
  python bench.py --steps 1 --warmup 1 --batch $B --no-extra --no-cpu-baseline > gpurun_out/ncu_scorer_b$B.log 2>&1
done
tail -12 gpurun_out/trace_b128_f0.txt gpurun_out/trace_b256_f0.txt gpurun_out/trace_b1024_f0.txt
