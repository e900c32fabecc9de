#!/bin/bash
# ncu evidence of the bench command on the GPU box (run through gpurun from the repository root; one GPU).  Numbers printed
# by runs under ncu are never bench values: only the launch list and the captures are kept.
R=${ROUND:-r2}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${R}_launches.csv $CMD > /dev/null 2>&1
# the first 1200 x 1200 hidden layer of FC-4 (second gemm_tc launch of the run) and the first conv_first launch
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 1 -c 1 -f -o gpurun_out/${R}_gemm_tc_full $CMD > /dev/null 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:conv_first_mma -s 2 -c 1 -f -o gpurun_out/${R}_conv_first_mma_full $CMD > /dev/null 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:merger_mma -s 1 -c 1 -f -o gpurun_out/${R}_merger_mma_full $CMD > /dev/null 2>&1
ls -la gpurun_out/${R}_launches.csv gpurun_out/${R}_*_full.ncu-rep
