#!/usr/bin/env bash
# compute-sanitizer over the small parity cases: memcheck (out-of-bounds / misaligned accesses, incl. the TMA bulk copies and
# cp.async gathers of the streaming kernels) and racecheck (shared-memory hazards in the warp-private rings).
set -u
mkdir -p gpurun_out
T="timeout -s KILL"
SEL='test_against_reference_dumps or test_jittered_unstructured_mesh_vs_oracle or test_monomial_basis_vs_oracle or test_device_built_teno_tables'
$T 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --target-processes all python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -c "Invalid\|misaligned" gpurun_out/sanitize_memcheck.log; tail -6 gpurun_out/sanitize_memcheck.log
$T 600 compute-sanitizer --tool memcheck --error-exitcode 99 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "partitioned" > gpurun_out/sanitize_memcheck_multi.log 2>&1
echo "memcheck multi rc=$?"; tail -4 gpurun_out/sanitize_memcheck_multi.log
$T 900 compute-sanitizer --tool racecheck --error-exitcode 99 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "test_against_reference_dumps and teno" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -c "hazard" gpurun_out/sanitize_racecheck.log; tail -8 gpurun_out/sanitize_racecheck.log
