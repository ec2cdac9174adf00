#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (ours + reference arm), ncu launch list + full capture of the top kernels.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $OUT/nproc.txt

echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1200 python -m pytest tests -q -m gpu >> $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log

echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?" >> $OUT/smoke.log
tail -3 $OUT/smoke.log

echo "== bench (reference arm)"
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?" >> $OUT/bench_ref.err
tail -c 900 $OUT/bench_ref.json

echo "== bench (ours)"
timeout 900 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
tail -c 3500 $OUT/bench.json; tail -3 $OUT/bench.err

echo "== ncu launch list (setup: first 400 launches)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_setup.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "rc=$?"; wc -l $OUT/launches_setup.csv

KERNELS='regex:grid_probe|pending_search|packet_search|factor_reduce'
echo "== ncu launch list of the timed region (the four launches of each linearize)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -s 40 -c 40 --csv --log-file $OUT/launches_timed.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches2.log 2>&1
echo "rc=$?"; wc -l $OUT/launches_timed.csv

echo "== ncu full capture of the linearize kernels (one pass over the 5 poses of the trajectory)"
timeout 1200 ncu --set full --clock-control none --import-source on -k "$KERNELS" -s 40 -c 20 -f -o $OUT/prof_linearize \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
echo "rc=$?"; ls -la $OUT | head -30
