#!/bin/bash
# fifth GPU pass: full single-GPU test suite; same-box A/B of the step kernels with evict-first ELL loads; e2e host sections
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
run() {  # name, env..., -- bench args
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 60 --warmup 5 --no-train --no-cpu-baseline $BARGS > gpurun_out/r2e_$name.json 2> gpurun_out/r2e_$name.err || echo "FAILED $name" >> gpurun_out/r2e_sweep.log
}
BARGS="--emulate-world 8"
run e8_classic_cs0 SML_STEP_KERNEL=classic SML_ELL_STREAM=0
run e8_classic_cs1 SML_STEP_KERNEL=classic SML_ELL_STREAM=1
run e8_p_s4_cs0 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4 SML_ELL_STREAM=0
run e8_p_s4_cs1 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4 SML_ELL_STREAM=1
run e8_p_c32s2_cs1 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=2 SML_PERSIST_STAGE_COLS=32 SML_ELL_STREAM=1
BARGS=""
run n1_classic_cs0 SML_STEP_KERNEL=classic SML_ELL_STREAM=0
run n1_classic_cs1 SML_STEP_KERNEL=classic SML_ELL_STREAM=1
run n1_p_s4_cs0 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4 SML_ELL_STREAM=0
run n1_p_s4_cs1 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4 SML_ELL_STREAM=1
run n1_p_s4_i1088_cs1 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=4 SML_ITEM_ROWS=1088 SML_ELL_STREAM=1
run n1_p_c32s2_cs1 SML_STEP_KERNEL=persist SML_PERSIST_STAGES=2 SML_PERSIST_STAGE_COLS=32 SML_ELL_STREAM=1
tail -5 gpurun_out/r2e_pytest.log
