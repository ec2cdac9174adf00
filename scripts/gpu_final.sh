#!/bin/bash
# Short GPU-box visit: parity tests, smoke, bench line, size sweep, launch list, reference arm (in that order of importance).
# Usage (from the repo root, under gpurun):  bash scripts/gpu_final.sh [tag]
TAG=${1:-r01ah}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> $OUT/nproc.txt
T0=$(date +%s)
LIMIT=${2:-840}   # seconds this visit may take in total; later stages are skipped when they no longer fit
fits() { [ $(( $(date +%s) - T0 + $1 )) -lt $LIMIT ]; }
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 540 python -m pytest tests -q -m gpu -x --durations=8 >> $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -14 $OUT/pytest.log; echo "t=$(( $(date +%s) - T0 ))s"
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "rc=$?" >> $OUT/smoke.log
tail -2 $OUT/smoke.log; echo "t=$(( $(date +%s) - T0 ))s"
echo "== bench (ours)"
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
tail -c 4500 $OUT/bench.json; tail -2 $OUT/bench.err; echo "t=$(( $(date +%s) - T0 ))s"
echo "== sweep"
fits 200 && timeout 200 python scripts/sweep.py --sizes 100000,300000,1000000,3000000,10000000 --reps 10 --out $OUT/sweep.json > $OUT/sweep.log 2>&1; echo "rc=$?" >> $OUT/sweep.log
tail -24 $OUT/sweep.log; echo "t=$(( $(date +%s) - T0 ))s"
echo "== ncu launch list of the timed region"
KERNELS='regex:grid_probe|pending_search|packet_search|factor_reduce'
fits 150 && timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KERNELS" -s 40 -c 40 --csv --log-file $OUT/launches_timed.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
echo "rc=$?"; wc -l $OUT/launches_timed.csv; echo "t=$(( $(date +%s) - T0 ))s"
echo "== bench (reference arm)"
fits 150 && timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "rc=$?" >> $OUT/bench_ref.err
tail -c 600 $OUT/bench_ref.json; echo "t=$(( $(date +%s) - T0 ))s"
