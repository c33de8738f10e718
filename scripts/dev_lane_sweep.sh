#!/bin/bash
# dev: rocket / grasp bench of the lane kernel at several instances-per-warp settings
V=${1:-noas}
export ALTRO_B200_LIB=gpurun_variants/libaltro_$V.so
for lpw in 4 8 16 32; do
  ALTRO_B200_LPW=$lpw python bench.py --workload rocket --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_rocket_${V}_lpw$lpw.json 2>> gpurun_out/r2f.err
done
ALTRO_B200_LPW=8 python bench.py --steps 50 --warmup 3 --workload grasp --no-cpu-baseline --no-e2e > gpurun_out/r2f_grasp_${V}_lpw8.json 2>> gpurun_out/r2f.err
