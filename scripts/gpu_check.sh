#!/bin/bash
# GPU-box visit: parity tests, smoke, bench line (short).  Usage (repo root, under gpurun): bash scripts/gpu_check.sh <tag> [pytest args]
TAG=${1:-r02a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $OUT/nproc.txt
T0=$(date +%s)
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 900 python -m pytest tests -q -m gpu --durations=10 ${@:2} >> $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -40 $OUT/pytest.log; echo "t=$(( $(date +%s) - T0 ))s"
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?" >> $OUT/smoke.log
tail -2 $OUT/smoke.log; echo "t=$(( $(date +%s) - T0 ))s"
echo "== bench (ours)"
timeout 400 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
tail -c 5000 $OUT/bench.json; tail -3 $OUT/bench.err; echo "t=$(( $(date +%s) - T0 ))s"
