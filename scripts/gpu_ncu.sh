#!/bin/bash
# ncu evidence for the three launches of sgb_linearize over one pass of the bench's pose schedule (T0 .. T4).
# Launch numbering (kernels matching the filter): rank 0's trajectory 5 x 3 = 15, the all-rank trajectory 5 x 3 = 15, warm-up 5 x 3 = 15,
# then the load roll: with SGB_BENCH_ROLL=5 its 15 launches are exactly one pass over the 5 poses -> -s 45 -c 15.
TAG=${1:-r02n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KERNELS='regex:grid_probe|packet_search|factor_reduce'
# (1) launch list of the timed region: device time of every launch (cold-cache, serialised: shares, not absolutes)
SGB_BENCH_ROLL=5 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -s 45 -c 30 --csv --log-file $OUT/launches_timed.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_launches.log 2>&1
echo "launch list rc=$?"; wc -l $OUT/launches_timed.csv
# (2) full capture of one pass over the poses
SGB_BENCH_ROLL=5 timeout 600 ncu --set full --clock-control none --import-source on -k "$KERNELS" -s 45 -c 15 -f -o $OUT/prof_linearize \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/ncu_full.log 2>&1
echo "full rc=$?"; tail -3 $OUT/ncu_full.log; ls -la $OUT
# (3) the bench line of the same build, for the record
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
