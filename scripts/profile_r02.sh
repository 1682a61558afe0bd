#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU): every capture only after the same command has exited 0 without ncu.
#   launch lists (cold-cache, serialised: compare shares) of the cfg2 and PNG device-resident steps and of the config-5 shape
#   ncu --set full of the dominant kernels: inflate_batch_kernel (cfg2), the lane-serial path + un-filter (PNG), block-split (cfg5)
set -u
O=gpurun_out
python scripts/bench_cfg2_dev.py 4096 > $O/p_cfg2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_cfg2.csv python scripts/bench_cfg2_dev.py 4096 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:inflate_batch_kernel -s 2 -c 1 -o $O/r02_inflate_cfg2 python scripts/bench_cfg2_dev.py 4096 > /dev/null 2>&1
python scripts/bench_png_shapes.py 1024 1024 2048 12 > $O/p_png.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_png_2048x1024.csv python scripts/bench_png_shapes.py 1024 1024 2048 12 > /dev/null 2>&1
python scripts/bench_png_shapes.py 1024 1024 512 12 > $O/p_png512.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fx_|png_unfilter|split_resolve|png_tasks" -s 11 -c 9 -o $O/r02_png_kernels python scripts/bench_png_shapes.py 1024 1024 512 12 > /dev/null 2>&1
python scripts/bench_png_shapes.py 8192 8192 8 1 > $O/p_png8192.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_png_8x8192.csv python scripts/bench_png_shapes.py 8192 8192 8 1 > /dev/null 2>&1
python scripts/bench_cfg5.py > $O/p_cfg5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_cfg5.csv python scripts/bench_cfg5.py > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"bs_count|bs_decode|bs_search" -s 3 -c 3 -o $O/r02_bsplit_kernels python scripts/bench_cfg5.py > /dev/null 2>&1
tail -n 1 $O/p_cfg2.log $O/p_png.log $O/p_png512.log $O/p_png8192.log $O/p_cfg5.log
ls -la $O/*.ncu-rep
