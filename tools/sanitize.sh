#!/bin/bash
# compute-sanitizer passes over the GPU test-suite (run on the GPU box, e.g. through gpurun):
#   memcheck  on every test except the full-size layers, racecheck on the kernels that do not use the async
#   proxy (pack, CUDA-core, XNOR) and, with known false positives on the bulk-copy ring (racecheck does not model
#   cp.async.bulk / mbarrier completion), the decode kernel -- racecheck does not model tcgen05 / TMA accesses.
# `tools/sanitize.sh decode` restricts both passes to a few small cases of the decode kernel and its index builder.
# The first `import torch` on a fresh box can outlast the sanitizer's attach timeout (it then never attaches and the
# target hangs): page torch in first and give the attach a generous --launch-timeout.
set -u
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
CS="compute-sanitizer --launch-timeout 300 --error-exitcode 9"
if [ "${1:-all}" = "pair" ]; then      # round 2: the pair decode kernel and the overlapped prefill expansion
  timeout ${2:-240} $CS --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "decode_pair_kernel_matches and dtype0 and (96-256 or 32-128 or 320-1152) and 3-True" \
    > gpurun_out/racecheck_pair.log 2>&1
  echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/racecheck_pair.log | tail -5
  timeout ${2:-240} $CS --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "(decode_pair_kernel_matches and dtype0 and (96-256 or 32-128 or 1000-384 or 320-1152)) or decode_pair_kernel_dense or decode_pair_kernel_symmetric or prefill_back_to_back or prefill_inside_cuda_graph or decode_kernel_activation" \
    > gpurun_out/memcheck_pair.log 2>&1
  echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_pair.log | tail -3
  exit 0
fi
if [ "${1:-all}" = "decode" ]; then
  timeout ${2:-200} $CS --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "(decode_kernel_matches and (100-70 or 96-160 or 264-1030) and dtype0) or (stream_layout_reproduces and 100-70 and dtype0) or (bireal_matches and (100-70 or 300-520) and xdtype0) or bireal_stream_k" \
    > gpurun_out/racecheck_decode.log 2>&1
  echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/racecheck_decode.log | tail -5
  timeout ${2:-200} $CS --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q \
    -k "(decode_kernel_matches and (100-70 or 300-520 or 256-512 or 2048-128) and dtype0) or (stream_layout_reproduces and dtype0) or decode_kernel_activation or (bireal_matches and (100-70 or 300-520 or 768-768)) or bireal_stream_k" \
    > gpurun_out/memcheck_decode.log 2>&1
  echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_decode.log | tail -3
  exit 0
fi
$CS --tool memcheck python -m pytest tests -m gpu -q --timeout 800 \
  -k "not full_size and not 4096 and not 11008 and not hf and not packed_checkpoint" > gpurun_out/memcheck_all.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_all.log | tail -3
$CS --tool racecheck python -m pytest tests/test_gpu_parity.py -q --timeout 400 \
  -k "(decode_kernel_matches and (100-70 or 264 or 96-160)) or (forward_matches and float32 and (100-70 or 96-160)) or (roundtrip and 100-70) or (bireal_matches and 100-70)" \
  > gpurun_out/racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/racecheck.log | tail -5
