#!/bin/bash
# Sweep the decode-kernel variants (one process each: the knobs are read once per process). Output: gpurun_out/decode_sweep.jsonl
mkdir -p gpurun_out
out=gpurun_out/decode_sweep.jsonl
: > $out
run() { env "$@" timeout 300 python tools/decode_bench.py --layers 4 >> $out 2>gpurun_out/decode_sweep.err || echo "{\"failed\": \"$*\"}" >> $out; }
run PBL_FORCE_KERNEL=2
run PBL_DK_CTAS=2 PBL_DK_OCC=3
run PBL_DK_CTAS=3 PBL_DK_OCC=3
run PBL_DK_CTAS=3 PBL_DK_OCC=3 PBL_PDL=0
run PBL_DK_CTAS=3 PBL_DK_OCC=4
run PBL_DK_CTAS=4 PBL_DK_OCC=4
run PBL_DK_CTAS=4 PBL_DK_OCC=4 PBL_PDL=0
cat $out
