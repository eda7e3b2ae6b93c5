#!/bin/bash
# Evidence run for profiles/ (round 2): plain bench, ncu launch list of the same command, ncu --set full of the hot kernels.
# The full capture is exported to CSV on the box and the .ncu-rep dropped: gpurun brings back at most 64 MiB.
set -x
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-learner"
$CMD > gpurun_out/r02_plain.log 2> gpurun_out/r02_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
python tools/r2_profile_target.py || exit 1
ncu --set full --clock-control none --import-source on -k regex:"env_advance|gather_|replay_sample" -o /tmp/r02_kernels -f python tools/r2_profile_target.py > gpurun_out/r02_ncu_full.log 2>&1
ncu -i /tmp/r02_kernels.ncu-rep --page raw --csv > gpurun_out/r02_kernels_raw.csv
ls -la /tmp/r02_kernels.ncu-rep gpurun_out/
tail -n 3 gpurun_out/r02_ncu_list.log gpurun_out/r02_ncu_full.log
