#!/usr/bin/env bash
# One-GPU visit: full parity suite, default bench line, the 16 M-cell cartesian mesh on one GPU (the strong-scaling base
# point), ncu launch list + full capture.  Output -> gpurun_out/
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
$T 1500 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
$T 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench.json
if [ "${REFARM:-1}" = 1 ]; then
$T 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; tail -c 600 gpurun_out/bench_reference.json
fi
if [ "${BIG:-1}" = 1 ]; then
$T 900 python bench.py --nx 8192 --ny 1024 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_16M_cart.json 2> gpurun_out/bench_16M_cart.err; echo "bench 16M rc=$?"; tail -c 900 gpurun_out/bench_16M_cart.json
fi
if [ "${NCU:-1}" = 1 ]; then
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
$T 900 ncu --set full --clock-control none --import-source on -k regex:'teno_stream|face_flux|gather_stage|cfl_kernel' -s 8 -c 8 -f -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out
