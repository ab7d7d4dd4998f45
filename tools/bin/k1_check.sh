#!/bin/bash
# stage-1 parity tests + a short device-resident bench (tuning loop for K1 / K2 / K3)
python -m pytest tests/test_gpu_stage1.py tests/test_gpu_batch.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-hamming --no-tracking "$@" 2>/dev/null | python tools/benchsum.py 2>/dev/null | head -7
