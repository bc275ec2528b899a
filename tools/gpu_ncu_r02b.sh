#!/bin/bash
# ncu evidence of the second half of round 2 (one GPU): full captures of the one-launch spin-up kernel and of the ring
# state-generation kernel, and a launch list of the bench step.  Each ncu command runs only after the same command exited 0
# without ncu (B200_PROFILING.md).
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-train --no-cpu-baseline"
$CMD > gpurun_out/ncu2_plain.log 2> gpurun_out/ncu2_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:k_sync_persist -s 1 -c 1 -f -o gpurun_out/prof_sync_persist_r02 $CMD > gpurun_out/ncu2_sync.log 2>&1
echo "sync capture rc=$?"
CMD2="python tools/bench_train.py --regions 144 --cols 600 --no-cublas"
$CMD2 > gpurun_out/ncu2_plain_t.log 2> gpurun_out/ncu2_plain_t.err &&
ncu --set full --clock-control none --import-source on -k regex:k_train_stategen_ring -s 1 -c 1 -f -o gpurun_out/prof_stategen_ring_r02 $CMD2 > gpurun_out/ncu2_sg.log 2>&1
echo "stategen capture rc=$?"
CMD3="python bench.py --steps 20 --warmup 3 --no-train --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 400 --csv --log-file gpurun_out/launches_r02b.csv $CMD3 > gpurun_out/ncu2_launches.log 2>&1
echo "launch list rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -4
