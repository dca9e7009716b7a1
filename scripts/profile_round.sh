#!/bin/bash
# Profiling pass of a round (run under gpurun): plain runs first, then the ncu launch list of the default bench command and one
# `--set full` capture each of the dense-refinement and the LK kernel.  Outputs land in gpurun_out/ (tag = $1).
set -u
TAG=${1:-r02}
OUT=gpurun_out
DPR="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-streams --frames 1024"
LK="python bench.py --workload lk --steps 2 --warmup 3 --frames 1024"
DEF="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-streams"
$DEF > $OUT/${TAG}_bench_launches_plain.json 2> $OUT/${TAG}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_bench_launches.csv $DEF > $OUT/${TAG}_bench_under_ncu.json 2> $OUT/${TAG}_ncu1.err
$DPR > $OUT/${TAG}_bench_dpr1024_plain.json 2>> $OUT/${TAG}_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:dpr_kernel -c 1 -f -o $OUT/${TAG}_dpr $DPR > $OUT/${TAG}_bench_dpr1024_under_ncu.json 2> $OUT/${TAG}_ncu2.err
$LK > $OUT/${TAG}_bench_lk1024_plain.json 2>> $OUT/${TAG}_plain.err &&
ncu --set full --clock-control none --import-source on -k regex:lk_kernel -c 1 -f -o $OUT/${TAG}_lk $LK > $OUT/${TAG}_bench_lk1024_under_ncu.json 2> $OUT/${TAG}_ncu3.err
echo profile pass done
