#!/bin/bash
# ncu --set full of the three launches of sgb_linearize at N points (default 10M): 2 calls at the identity pose, 2 at the converged pose.
# usage (under gpurun): bash scripts/gpu_ncu_size.sh <tag> [N]
TAG=${1:-r02size}; N=${2:-10000000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python scripts/linearize_at_size.py $N 3 > $OUT/plain.log 2>&1; echo "plain rc=$?"; cat $OUT/plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:grid_probe|packet_search|factor_reduce' -f -o $OUT/prof_size \
    python scripts/linearize_at_size.py $N 2 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/ncu.log; ls -la $OUT
