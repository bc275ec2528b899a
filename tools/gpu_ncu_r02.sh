#!/bin/bash
# ncu evidence of round 2 (one GPU): launch list of the bench step, full captures of the persistent step kernel and of the
# update ring kernel.  Each ncu command runs only after the same command exited 0 without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 3 --no-train --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2> gpurun_out/ncu_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline"
$CMD2 > gpurun_out/ncu_plain2.log 2> gpurun_out/ncu_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:k_step_persist -s 14 -c 2 -f -o gpurun_out/prof_step_persist_r02 $CMD2 > gpurun_out/ncu_step.log 2>&1
echo "step capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_update_ring -s 6 -c 2 -f -o gpurun_out/prof_update_ring_r02 $CMD2 > gpurun_out/ncu_update.log 2>&1
echo "update capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:k_readout_finish -s 14 -c 2 -f -o gpurun_out/prof_finish_r02 $CMD2 > gpurun_out/ncu_finish.log 2>&1
echo "finish capture rc=$?"
ls -la gpurun_out/*.ncu-rep
