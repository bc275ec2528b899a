#!/bin/bash
# first GPU pass of round 2: host facts, the single-GPU test suite, a short bench line
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi -L; } > gpurun_out/r2a_host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?" >> gpurun_out/r2a_bench.err
SML_STEP_KERNEL=classic timeout 600 python bench.py --steps 20 --warmup 5 --no-train --no-cpu-baseline > gpurun_out/r2a_bench_classic.json 2> gpurun_out/r2a_bench_classic.err
echo "bench rc=$?" >> gpurun_out/r2a_bench_classic.err
tail -5 gpurun_out/r2a_pytest.log
