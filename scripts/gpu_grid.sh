#!/bin/bash
# Grid front-end tuning round: parity tests, a sweep of the cell size, per-launch times and one full ncu capture.
TAG=${1:-r01n}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1200 python -m pytest tests -q -m gpu -x >> $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -8 $OUT/pytest.log
show() {
python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2 value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "warm", round(d["value_l2_warm"],1), "launches", d["gpu_launches"], "iters", d["pose_error_vs_gt"]["gn_iterations"], "parity", d.get("parity_rel_H"))
except Exception as e:
    print("fail", e); print(open("$1".replace(".json",".err")).read()[-1500:])
PY
}
for c in 0 1.5 2 3 4; do
if [ "$c" = "0" ]; then export SGB_GRID=0; else export SGB_GRID=1 SGB_GRID_CELL=$c; fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_cell$c.json 2> $OUT/bench_cell$c.err
show $OUT/bench_cell$c.json "cell=$c"
done
export SGB_GRID=1
for c in 2 3; do
echo "== pending counts, cell=$c"
SGB_GRID_CELL=$c SGB_DEBUG_PENDING=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "grid front end" | sort | uniq -c | sort -rn | head -12
echo "== ncu launch list of the timed region (cell=$c)"
SGB_GRID_CELL=$c timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"packet_search|factor_reduce|grid_probe|pending_search" -s 40 -c 48 --csv --log-file $OUT/launches_cell$c.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("$OUT/launches_cell$c.csv")))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:30], float(r[vi].replace(',',''))/1e3) for r in rows[start:] if len(r)>vi]
for k,v in seq[:16]: print(f"{k:32s} {v:8.1f} us")
PY
done
echo "== ncu full: probe + pending (cell=3)"
SGB_GRID_CELL=3 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"grid_probe|pending_search" -s 24 -c 2 -o $OUT/grid_full -f \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls -la $OUT
