#!/bin/bash
# third GPU pass: full single-GPU test suite, persistent-kernel sweep with the unpadded single-copy tiles, update-kernel A/B
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 60 --warmup 5 --no-train --no-cpu-baseline $BARGS > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err || echo "FAILED $name" >> gpurun_out/r2c_sweep.log
}
BARGS="--emulate-world 8"
run e8_classic_sx SML_STEP_KERNEL=classic SML_UPDATE_KERNEL=sx
run e8_p272_s5 SML_STEP_KERNEL=persist
run e8_p272_s4 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4
run e8_p272_c32s3 SML_STEP_KERNEL=persist SML_PERSIST_STAGE_COLS=32 SML_PERSIST_STAGES=3
run e8_p128_s5 SML_STEP_KERNEL=persist SML_PART_ROWS=128
run e8_p272_i1088 SML_STEP_KERNEL=persist SML_ITEM_ROWS=1088
run e8_p272_pad SML_STEP_KERNEL=persist SML_PERSIST_PAD=1
run e8_ring4 SML_STEP_KERNEL=persist SML_UPDATE_STAGES=4
run e8_ringsplit2 SML_STEP_KERNEL=persist SML_UPDATE_SPLIT=2
BARGS=""
run n1_classic_sx SML_STEP_KERNEL=classic SML_UPDATE_KERNEL=sx
run n1_p272_s5 SML_STEP_KERNEL=persist
run n1_p272_c32s3 SML_STEP_KERNEL=persist SML_PERSIST_STAGE_COLS=32 SML_PERSIST_STAGES=3
run n1_ringsplit2 SML_STEP_KERNEL=persist SML_UPDATE_SPLIT=2
tail -5 gpurun_out/r2c_pytest.log
