#!/bin/bash
# second GPU pass: full single-GPU test suite, then the persistent-kernel parameter sweep (emulated 8-rank shard and N=1)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 60 --warmup 5 --no-train --no-cpu-baseline $BARGS > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err || echo "FAILED $name" >> gpurun_out/r2b_sweep.log
}
BARGS="--emulate-world 8"
run e8_classic SML_STEP_KERNEL=classic
run e8_p272_s5 SML_STEP_KERNEL=persist
run e8_p272_s4 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4
run e8_p272_s6 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=6
run e8_p128_s5 SML_STEP_KERNEL=persist SML_PART_ROWS=128
run e8_p192_i576 SML_STEP_KERNEL=persist SML_PART_ROWS=192 SML_ITEM_ROWS=576
run e8_p272_i1088 SML_STEP_KERNEL=persist SML_ITEM_ROWS=1088
run e8_p272_nostag SML_STEP_KERNEL=persist SML_PERSIST_STAGGER=0
run e8_p272_c32 SML_STEP_KERNEL=persist SML_PERSIST_STAGE_COLS=32 SML_PERSIST_STAGES=3
BARGS=""
run n1_classic SML_STEP_KERNEL=classic
run n1_p272_s5 SML_STEP_KERNEL=persist
run n1_p128_s5 SML_STEP_KERNEL=persist SML_PART_ROWS=128
run n1_p272_s6 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=6
tail -5 gpurun_out/r2b_pytest.log
