#!/bin/bash
# Full ncu capture of the four launches of sgb_linearize over one pass of the bench's poses (T0 .. T4).
# Launch numbering (kernels matching the filter): Gauss-Newton trajectory 5 x 4 = 20, warm-up 3 x 4 = 12, then the load roll:
# with SGB_BENCH_ROLL=5 its 20 launches are exactly one pass over the 5 poses -> -s 32 -c 20.
TAG=${1:-r01am}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KERNELS='regex:grid_probe|pending_search|packet_search|factor_reduce'
SGB_BENCH_ROLL=5 timeout 300 ncu --set full --clock-control none --import-source on -k "$KERNELS" -s 32 -c 20 -f -o $OUT/prof_linearize \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "rc=$?"; tail -3 $OUT/ncu_full.log; ls -la $OUT
