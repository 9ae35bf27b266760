#!/bin/bash
# ncu evidence for the decode kernel: launch list (old skinny vs decode kernel) + one full capture per layer shape.
mkdir -p gpurun_out
PBL_FORCE_KERNEL=2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'skinny|decode_mma' -c 40 --csv \
    --log-file gpurun_out/launches_skinny.csv python tools/decode_bench.py --layers 1 --reps 1 > gpurun_out/prof_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'skinny|decode_mma' -c 40 --csv \
    --log-file gpurun_out/launches_decode.csv python tools/decode_bench.py --layers 1 --reps 1 > gpurun_out/prof_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_mma -s 10 -c 7 -o gpurun_out/decode_full -f \
    python tools/decode_bench.py --layers 1 --reps 1 > gpurun_out/prof_c.log 2>&1
tail -3 gpurun_out/prof_c.log
ls -la gpurun_out/
