#!/usr/bin/env bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full capture of the hot kernels.  Output -> gpurun_out/
# Every step runs under its own hard timeout so that a hung kernel cannot take the box down with it.
set -u
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; nvidia-smi -L >> gpurun_out/host.txt; free -g >> gpurun_out/host.txt
T="timeout -s KILL"
$T 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
if [ "${BENCH:-1}" = 1 ]; then
$T 600 python bench.py --steps 20 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ -n "${VARIANT:-}" ]; then
MLB_STREAM_GATHER=$VARIANT $T 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e ${BENCH_ARGS:-} > gpurun_out/bench_$VARIANT.json 2> gpurun_out/bench_$VARIANT.err; echo "bench $VARIANT rc=$?"
tail -c 600 gpurun_out/bench_$VARIANT.json
fi
if [ "${STRICT:-1}" = 1 ]; then
$T 600 python bench.py --steps 20 --warmup 3 --fp strict --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/bench_strict.json 2> gpurun_out/bench_strict.err; echo "bench strict rc=$?"
tail -c 1000 gpurun_out/bench_strict.json
fi
fi
if [ "${NCU:-1}" = 1 ]; then
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/ncu_launches.log 2>&1
$T 900 ncu --set full --clock-control none --import-source on -k regex:'teno_|face_flux|gather_stage|cfl_kernel' -s 8 -c 7 -f -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
fi
