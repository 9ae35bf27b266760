#!/bin/bash
# compute-sanitizer passes over the GPU test-suite (run on the GPU box, e.g. through gpurun):
#   memcheck  on every test except the full-size layers, racecheck on the kernels that do not use the async
#   proxy (pack, CUDA-core, mma.sync skinny, XNOR) -- racecheck does not model tcgen05 / TMA accesses.
set -u
mkdir -p gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q --timeout 800 \
  -k "not full_size and not 4096 and not 11008 and not hf and not packed_checkpoint" > gpurun_out/memcheck_all.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_all.log | tail -3
compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q --timeout 400 \
  -k "(skinny and (100-70 or 264 or 96-160)) or (forward_matches and float32 and (100-70 or 96-160)) or (roundtrip and 100-70) or (bireal_matches and 100-70)" \
  > gpurun_out/racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/racecheck.log | tail -5
