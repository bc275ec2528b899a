timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu5.log 2>&1; echo pytest rc=$?; tail -4 gpurun_out/pytest_gpu5.log
for r in 4 8; do SML_UPDATE_RPT=$r timeout 300 python tools/sweep.py --m 6000 --deg 6 24 > gpurun_out/sweep_upd$r.jsonl 2>gpurun_out/sweep.err; python - gpurun_out/sweep_upd$r.jsonl rpt=$r <<'PY'
import json,sys
for l in open(sys.argv[1]):
    r=json.loads(l); print(sys.argv[2], r["m"],r["degree"],"update_ms",round(r["update_ms"],4),"GB/s",round(r["update_GBs"]),"frac",round(r["update_frac"],3))
PY
done
for lpt in 0 1 0 1; do unset SML_NO_LPT; [ $lpt = 1 ] && export SML_NO_LPT=1; timeout 200 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-train > gpurun_out/ab_lpt.log 2>gpurun_out/ab.err; echo "no_lpt=$lpt $(grep -o "\"ms_per_step\": [0-9.]*\|\"kernel_ms_per_launch\": [0-9.]*" gpurun_out/ab_lpt.log | tr "\n" " ")"; done
unset SML_NO_LPT
timeout 900 python tools/sweep.py --md gpurun_out/sweep_r01.md > gpurun_out/sweep_r01.jsonl 2> gpurun_out/sweep.err; echo sweep rc=$?; cat gpurun_out/sweep_r01.md
