#!/bin/bash
# Round-2 profiles: launch list and one full ncu capture of the timed closed-loop launch, rocket and quadruped.
set -x
CMD_R="python bench.py --workload rocket --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
CMD_Q="python bench.py --workload quadruped --steps 20 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD_R > gpurun_out/r2p_rocket_plain.json 2> gpurun_out/r2p_rocket_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_rocket.csv $CMD_R > gpurun_out/r2p_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:altro_solve_kernel -s 3 -c 1 -f -o gpurun_out/r2_rocket_run $CMD_R > gpurun_out/r2p_ncu2.log 2>&1
$CMD_Q > gpurun_out/r2p_quad_plain.json 2> gpurun_out/r2p_quad_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:altro_solve_kernel -s 2 -c 1 -f -o gpurun_out/r2_quadruped_run $CMD_Q > gpurun_out/r2p_ncu3.log 2>&1
tail -2 gpurun_out/r2p_ncu1.log gpurun_out/r2p_ncu2.log gpurun_out/r2p_ncu3.log
