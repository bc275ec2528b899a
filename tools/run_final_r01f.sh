# round-1 closing GPU pass: LU route after the buffer-guard refactor, symbols, default bench with the update-only leg
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_train_gpu.py tests/test_config1_gpu.py tests/test_capi_symbols.py -x -q > gpurun_out/pytest_f2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_f2.log
timeout 600 python bench.py > gpurun_out/bench_r01f.log 2> gpurun_out/bench_r01f.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r01f.log").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["cpu_baseline"]["value"],
      d["train"]["value"], d["train"]["solve_ms_per_region"], d["train"]["stategen_ms"])
print(d["update_roofline"])
PY
