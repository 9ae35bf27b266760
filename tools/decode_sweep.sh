#!/bin/bash
# Sweep the decode-kernel variants (one process each: the knobs are read once per process). Output: gpurun_out/decode_sweep.jsonl
mkdir -p gpurun_out
out=gpurun_out/decode_sweep.jsonl
: > $out
run() { env "$@" timeout 300 python tools/decode_bench.py --layers 4 >> $out 2>gpurun_out/decode_sweep.err || echo "{\"failed\": \"$*\"}" >> $out; }
for v in "$@"; do run $v; done
cat $out
