#!/bin/bash
# Round-2 profiling pass (one gpurun call): launch lists and `ncu --set full` captures of the training-step MLP kernels
# and of the render forward kernel.  Every ncu command follows a plain run of the same command that exited 0.
set -x
T="python bench.py --profile train --steps 4"
R="python bench.py --profile render --steps 1 --warmup 3"
$T > gpurun_out/plain_train.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 200 --csv --log-file gpurun_out/launches_r02_train.csv $T > /dev/null 2>&1
$R > gpurun_out/plain_render.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_r02_render.csv $R > /dev/null 2>&1
$T > gpurun_out/plain_train2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:mlp_fwd_kernel|mlp_bwd_dgrad|mlp_bwd_wgrad" -s 60 -c 6 -f -o gpurun_out/mlp_train_r02 $T > /dev/null 2>&1
$R > gpurun_out/plain_render2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_fwd_kernel -s 6 -c 2 -f -o gpurun_out/mlp_fwd_r02 $R > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r02_*.csv
