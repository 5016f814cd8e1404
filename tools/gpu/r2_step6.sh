#!/bin/bash
# 2-GPU box: new service test (frontend fixture), lanes in encode_rows, 2-rank parity, hybrid breakdown at 2 ranks, encode profile
mkdir -p gpurun_out; rm -f gpurun_out/summary.txt
for f in test_gpu_service test_gpu_towers; do
  timeout 900 python -m pytest tests/$f.py -q -m gpu -x --timeout=600 -s > gpurun_out/$f.log 2>&1
  echo "$f exit $? $(tail -1 gpurun_out/$f.log)" >> gpurun_out/summary.txt
done
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu --timeout=800 -s > gpurun_out/test_gpu_multi_n2.log 2>&1
echo "test_gpu_multi exit $? $(tail -1 gpurun_out/test_gpu_multi_n2.log)" >> gpurun_out/summary.txt
for cfg in "8841823 4096" "8841823 256" "2000000 4096"; do set -- $cfg
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 \
    tools/hybrid_breakdown_n.py $1 $2 > gpurun_out/hybrid_breakdown_n2_$1_$2.txt 2>&1
tail -1 gpurun_out/hybrid_breakdown_n2_$1_$2.txt >> gpurun_out/summary.txt
done
timeout 600 python tools/encode_rows_profile.py 400000 > gpurun_out/encode_rows_profile.txt 2>&1
grep lanes gpurun_out/encode_rows_profile.txt >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
