#!/usr/bin/env bash
# Multi-GPU visit (gpurun --gpus N): NCCL parity test + weak-scaling bench lines for 1..N GPUs.  Output -> gpurun_out/
set -u
mkdir -p gpurun_out
N=${NGPU:-2}
T="timeout -s KILL"
nvidia-smi -L > gpurun_out/multi_host.txt; nproc >> gpurun_out/multi_host.txt
$T 600 python -m pytest ${PYTEST_TARGET:-tests/test_gpu_multi.py} -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
for n in ${NLIST:-$N}; do
  if [ "$n" = 1 ]; then
    $T 600 python bench.py --gpus 1 --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
  else
    $T 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  fi
  echo "bench n=$n rc=$?"; tail -c 1800 gpurun_out/bench_n$n.json; tail -3 gpurun_out/bench_n$n.err
done
