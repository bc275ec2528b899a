#!/bin/bash
# multi-GPU pass: N = number of visible GPUs.  tests/test_multigpu.py (engine exchange vs NCCL vs oracle, C++ replay at 2 ranks),
# then bench.py at N (and at 2 / 4 on the way up when N = 8)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > gpurun_out/mg_topo.txt 2>&1
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 1500 python -m pytest tests/test_multigpu.py -m gpu -q -rs > gpurun_out/mg_pytest_n$N.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/mg_pytest_n$N.log
fi
for n in 2 4 8; do
  if [ $n -le $N ]; then
    port=$((29500 + n))
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
        bench.py --gpus $n --steps ${STEPS:-40} --warmup 5 $EXTRA > gpurun_out/mg_bench_n$n.json 2> gpurun_out/mg_bench_n$n.err
    echo "bench n=$n rc=$?" >> gpurun_out/mg_pytest_n$N.log
  fi
done
tail -6 gpurun_out/mg_pytest_n$N.log
