# round-1 last GPU pass on the final code: full GPU suite, smoke(), default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r01e.log 2>&1; echo "pytest gpu rc=$?"; tail -3 gpurun_out/pytest_gpu_r01e.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01e.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r01e.log
timeout 600 python bench.py > gpurun_out/bench_r01e.log 2> gpurun_out/bench_r01e.err; echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r01e.log").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["cpu_baseline"]["value"],
      d["train"]["value"], d["train"]["solve_ms_per_region"], d["train"]["stategen_ms"])
PY
