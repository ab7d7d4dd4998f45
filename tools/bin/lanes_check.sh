#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
F="--steps 3 --warmup 3 --no-cpu-baseline --no-hamming --no-tracking"
for l in 1 2; do for wi in 384 768 1536; do echo "LANES $l WI $wi"; PSLAM_LANES=$l python bench.py $F --work-images $wi 2>/dev/null | python tools/benchsum.py 2>/dev/null | head -2; done; done
echo BH188; PSLAM_K1_BH=188 python bench.py $F 2>/dev/null | python tools/benchsum.py 2>/dev/null | head -2
