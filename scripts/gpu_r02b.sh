#!/usr/bin/env bash
# Round 2, visit B (2 GPUs): partitioned-execution tests incl. the native NCCL driver, the 2-GPU bench line with strong records.
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
nvidia-smi -L
$T 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -rs -x --durations=8 -k "${KEXPR:-multi or partition or local or nccl or asynch or reference_dumps or riemann}" > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -25 gpurun_out/pytest_multi.log
N=${NGPU:-2}
MLB_PREP_TIMING=1 $T 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"
tail -c 3000 gpurun_out/bench_n$N.json; tail -30 gpurun_out/bench_n$N.err
ls -la gpurun_out
