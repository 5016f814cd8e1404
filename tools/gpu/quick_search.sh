#!/bin/bash
# quick search check: parity tests + bench on the full corpus and on 1/2, 1/4, 1/8 shards
timeout 400 python -m pytest tests/test_gpu_search.py tests/test_gpu_hybrid.py -q -m gpu -x --timeout=300 2>&1 | tail -1
for cfg in "8841823 128" "4420912 128" "2210456 128" "1105228 128" "8841823 16"; do set -- $cfg
timeout 300 python bench.py --steps 20 --warmup 3 --docs $1 --batch $2 --no-extra --no-cpu-baseline | grep -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*' | tr '\n' ' '; echo " <- $cfg"; done
