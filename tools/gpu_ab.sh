#!/bin/bash
# A/B of column tile widths on one B200 (run under gpurun)
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for w in 4 2; do XFB_COL_W=$w python bench.py --grid 4096 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 | python tools/benchsum.py "W=$w"; done
for w in 2 1; do XFB_COL_W=$w python bench.py --grid 8192 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 | python tools/benchsum.py "W=$w"; done
python bench.py --grid 2048 --steps 50 --warmup 3 --no-cpu-baseline --e2e-steps 3 | python tools/benchsum.py
