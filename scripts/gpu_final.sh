#!/usr/bin/env bash
# Final validation of a round on one GPU: full parity suite, smoke(), default bench line, reference arm, STRICT-mode bench,
# ncu launch list + full capture.  Output -> gpurun_out/
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
$T 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
$T 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
$T 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/bench.json
$T 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference arm rc=$?"; tail -c 500 gpurun_out/bench_reference.json
$T 600 python bench.py --fp strict --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_strict.json 2> gpurun_out/bench_strict.err; echo "strict rc=$?"; tail -c 600 gpurun_out/bench_strict.json
$T 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
$T 900 ncu --set full --clock-control none --import-source on -k regex:'teno_stream|face_flux|gather_stage|cfl_kernel' -s 8 -c 8 -f -o gpurun_out/prof \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
$T 600 ncu --set full --clock-control none -k regex:'teno_strict' -s 3 -c 1 -f -o gpurun_out/prof_strict \
    python bench.py --fp strict --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_strict.log 2>&1
ls -la gpurun_out
