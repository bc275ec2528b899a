#!/bin/bash
# One parameterised GPU pass (replaces the per-letter scripts of rounds 1 and 2).  Run under gpurun:
#     gpurun --timeout 3000 -- 'bash tools/gpu_pass.sh TAG [--tests] [--sweep FILE] [--steps N]'
# --tests        the whole -m gpu suite first (gpurun_out/TAG_pytest.log)
# --sweep FILE   one bench.py run per non-comment line of FILE:   name | ENV=val ENV=val ... | bench.py arguments
#                -> gpurun_out/TAG_name.json / .err.  Example line:
#                    e8_persist | SML_STEP_KERNEL=persist | --emulate-world 8
# Every measured alternative of the engine is an environment switch, so an A/B on ONE box -- the only
# comparison that survives the 3 % box-to-box spread -- is a sweep file.
set -u
TAG=$1; shift
TESTS=0; SWEEP=""; STEPS=60
while [ $# -gt 0 ]; do
  case "$1" in
    --tests) TESTS=1 ;;
    --sweep) SWEEP=$2; shift ;;
    --steps) STEPS=$2; shift ;;
  esac
  shift
done
mkdir -p gpurun_out
{ nproc; free -g | head -2; nvidia-smi -L; } > gpurun_out/${TAG}_host.txt 2>&1
if [ $TESTS = 1 ]; then
  timeout 2400 python -m pytest tests -m gpu -q -rs > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
  tail -4 gpurun_out/${TAG}_pytest.log
fi
if [ -n "$SWEEP" ]; then
  grep -v '^\s*#' "$SWEEP" | grep '|' | while IFS='|' read -r name envs bargs; do
    name=$(echo $name); 
    env $envs timeout 600 python bench.py --steps $STEPS --warmup 5 $bargs > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err \
      || echo "FAILED $name" >> gpurun_out/${TAG}_sweep.log
  done
fi
