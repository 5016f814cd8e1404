#!/bin/bash
# Round-2 ncu evidence (1 GPU).  Every command is first run without ncu; then launch lists (gpu__time_duration.sum,
# --clock-control none) of the default bench command, the encode and the training micro-benchmarks, and --set full
# captures of the dominant kernels in their round-2 form.
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
timeout 300 python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 || exit 1
timeout 300 python tools/encode_bench.py 7680 7680 > gpurun_out/plain_encode.log 2>&1 || exit 1
timeout 300 python tools/train_bench.py > gpurun_out/plain_train.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_launches_search_b128.csv \
  python bench.py --steps 3 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_launches_search_b128_shard8.csv \
  python bench.py --steps 3 --warmup 3 --docs 1105228 --no-extra --no-cpu-baseline > gpurun_out/ncu1b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_launches_search_1M_b256.csv \
  python bench.py --steps 3 --warmup 3 --docs 1000000 --batch 256 --no-extra --no-cpu-baseline > gpurun_out/ncu1c.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_launches_search_1M_b1.csv \
  python bench.py --steps 3 --warmup 3 --docs 1000000 --batch 1 --no-extra --no-cpu-baseline > gpurun_out/ncu1d.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_encode.csv \
  python tools/encode_bench.py 7680 7680 > gpurun_out/ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_train.csv \
  python tools/train_bench.py > gpurun_out/ncu2b.log 2>&1
# full captures: the 8.84 M-document scan at B = 128 (the roofline kernel of the default line), the 1/8-shard fused launch,
# the CTA-pair scan at B = 256 on 1 M documents, both projection GEMMs, the recurrence
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:score_topk_mma_kernel -s 5 -c 1 \
  -o gpurun_out/r2_prof_scorer_b128_8p8M -f python bench.py --steps 2 --warmup 1 --no-extra --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:score_topk_mma_kernel -s 3 -c 1 \
  -o gpurun_out/r2_prof_scorer_b128_shard8 -f python bench.py --steps 2 --warmup 1 --docs 1105228 --no-extra --no-cpu-baseline > gpurun_out/ncu3b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:score_topk_mma_kernel -s 3 -c 1 \
  -o gpurun_out/r2_prof_scorer_b256_1M -f python bench.py --steps 2 --warmup 1 --docs 1000000 --batch 256 --no-extra --no-cpu-baseline > gpurun_out/ncu3c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_bias_pair_kernel -s 4 -c 2 \
  -o gpurun_out/r2_prof_gemm_pair -f python tools/encode_bench.py 7680 7680 > gpurun_out/ncu5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:gru_fwd_tc_kernel -s 2 -c 1 \
  -o gpurun_out/r2_prof_gru_fwd_tc -f python tools/encode_bench.py 7680 7680 > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep | awk '{print $5, $9}'
