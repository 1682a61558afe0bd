#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU): every capture only after the same command has exited 0 without ncu.
#   launch lists (cold-cache, serialised: compare shares) of the cfg2 and PNG device-resident steps and of the config-5 shape
#   ncu --set full of the dominant kernels: inflate_batch_kernel (cfg2), the lane-serial path + resolve + un-filter (PNG),
#   block-split (cfg5)
# usage: profile_r02.sh a|b   (two calls: what a call writes under gpurun_out/ must stay below 64 MiB to be copied back)
set -u
O=gpurun_out
T="timeout 300"
if [ "${1:-a}" = a ]; then
$T python scripts/bench_cfg2_dev.py 4096 > $O/p_cfg2.log 2>&1 &&
$T ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_cfg2.csv python scripts/bench_cfg2_dev.py 4096 > /dev/null 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:inflate_batch_kernel -s 2 -c 1 -o $O/r02_inflate_cfg2 python scripts/bench_cfg2_dev.py 4096 > /dev/null 2>&1
$T python scripts/bench_cfg5.py > $O/p_cfg5.log 2>&1 &&
$T ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_cfg5.csv python scripts/bench_cfg5.py > /dev/null 2>&1
$T ncu --set full --clock-control none --import-source on -k regex:"bs_count|bs_decode|bs_search" -s 3 -c 3 -o $O/r02_bsplit_kernels python scripts/bench_cfg5.py > /dev/null 2>&1
tail -n 1 $O/p_cfg2.log $O/p_cfg5.log
else
PNG_FILT=mix $T python scripts/bench_png_shapes.py 1024 1024 2048 12 > $O/p_png.log 2>&1 &&
PNG_FILT=mix $T ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_png_2048x1024.csv python scripts/bench_png_shapes.py 1024 1024 2048 12 > /dev/null 2>&1
PNG_FILT=mix $T python scripts/bench_png_shapes.py 1024 1024 512 12 > $O/p_png512.log 2>&1 &&
PNG_FILT=mix $T ncu --set full --clock-control none --import-source on -k regex:"fx_head|fx_sizes|fx_tokens|fx_expand|png_unfilter|split_resolve|png_tasks" -s 8 -c 8 -o $O/r02_png_kernels python scripts/bench_png_shapes.py 1024 1024 512 12 > /dev/null 2>&1
$T python scripts/bench_png_shapes.py 8192 8192 8 1 > $O/p_png8192.log 2>&1 &&
$T ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_png_8x8192.csv python scripts/bench_png_shapes.py 8192 8192 8 1 > /dev/null 2>&1
tail -n 1 $O/p_png.log $O/p_png512.log $O/p_png8192.log
fi
ls -la $O/r02_*.ncu-rep
du -sm $O
