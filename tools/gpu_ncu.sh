#!/bin/bash
# one ncu --set full capture of the two stepper kernels + the launch list (run under gpurun)
GRID=${1:-4096}
CMD="python bench.py --grid $GRID --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3"
$CMD > gpurun_out/plain_$GRID.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"col_kernel|row_kernel" -s 14 -c 2 -o gpurun_out/prof_$GRID $CMD > gpurun_out/ncu_$GRID.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$GRID.csv $CMD > gpurun_out/ncu2_$GRID.log 2>&1
tail -2 gpurun_out/ncu_$GRID.log
