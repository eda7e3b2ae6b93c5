#!/bin/bash
# Evidence run for profiles/: plain bench, ncu launch list of the same command, ncu --set full of the hot kernels.
set -x
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r01_plain.log 2> gpurun_out/r01_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/r01_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:env_advance -s 30 -c 2 -o gpurun_out/r01_env_advance -f $CMD > gpurun_out/r01_ncu_a.log 2>&1
# bench.py's replay section launches, per layout, 63 single-minibatch gathers of 32 and 63 of 512 before the 8,192-transition calls
ncu --set full --clock-control none --import-source on -k regex:gather_u8 -s 130 -c 2 -o gpurun_out/r01_replay_u8 -f $CMD > gpurun_out/r01_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gather_f32 -s 130 -c 2 -o gpurun_out/r01_replay_f32 -f $CMD > gpurun_out/r01_ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:replay_sample -s 270 -c 2 -o gpurun_out/r01_replay_sample -f $CMD > gpurun_out/r01_ncu_d.log 2>&1
tail -n 2 gpurun_out/r01_ncu_a.log gpurun_out/r01_ncu_b.log gpurun_out/r01_ncu_c.log gpurun_out/r01_ncu_d.log
