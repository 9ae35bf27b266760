#!/bin/bash
# ncu evidence for the decode kernel: launch list (old skinny vs decode kernel) + full captures (one per layer shape).
mkdir -p gpurun_out
export PBL_DK_CTAS=${PBL_DK_CTAS:-3}
if [ "$1" != "quick" ]; then
PBL_FORCE_KERNEL=2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'skinny|decode_mma' -c 40 --csv \
    --log-file gpurun_out/launches_skinny.csv python tools/decode_bench.py --layers 1 --reps 1 > gpurun_out/prof_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'skinny|decode_mma' -c 40 --csv \
    --log-file gpurun_out/launches_decode.csv python tools/decode_bench.py --layers 1 --reps 1 > gpurun_out/prof_b.log 2>&1
fi
ncu --set full --clock-control none --import-source on --warp-sampling-interval 0 -k regex:decode_mma -s 13 -c 3 -o gpurun_out/decode_full -f \
    python tools/decode_bench.py --layers 1 --reps 1 > gpurun_out/prof_c.log 2>&1
tail -3 gpurun_out/prof_c.log
