#!/bin/bash
# fourth GPU pass: full single-GPU test suite, persistent-kernel ring depth at 144 / 288 / 576 / 1152 regions per GPU,
# update-kernel A/B (ring with large tiles vs sx)
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 60 --warmup 5 --no-train --no-cpu-baseline $BARGS > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err || echo "FAILED $name" >> gpurun_out/r2d_sweep.log
}
BARGS="--emulate-world 8"
run e8_classic_sx SML_STEP_KERNEL=classic SML_UPDATE_KERNEL=sx
run e8_p_s3 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=3
run e8_p_s4 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4
run e8_p_s5 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=5
run e8_p_s4_p128 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4 SML_PART_ROWS=128
run e8_p_s4_ring3 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4 SML_UPDATE_STAGES=3
BARGS="--emulate-world 4"
run e4_classic SML_STEP_KERNEL=classic SML_UPDATE_KERNEL=sx
run e4_p_s4 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4
BARGS="--emulate-world 2"
run e2_classic SML_STEP_KERNEL=classic SML_UPDATE_KERNEL=sx
run e2_p_s4 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4
BARGS=""
run n1_p_s3 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=3
run n1_p_s4 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4
run n1_p_s4_ring3 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4 SML_UPDATE_STAGES=3
tail -5 gpurun_out/r2d_pytest.log
