#!/bin/bash
TAG=${1:-r01m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1200 python -m pytest tests -q -m gpu >> $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -12 $OUT/pytest.log
show() {
python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2 value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "warm", round(d["value_l2_warm"],1), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "iters", d["pose_error_vs_gt"]["gn_iterations"])
except Exception as e:
    print("fail", e); print(open("$1".replace(".json",".err")).read()[-1500:])
PY
}
for g in 1 0; do
SGB_GRID=$g timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_grid$g.json 2> $OUT/bench_grid$g.err
show $OUT/bench_grid$g.json "grid=$g"
done
echo "== ncu launch list of the timed region (grid on)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"packet_search|factor_reduce|grid_probe" -s 30 -c 60 --csv --log-file $OUT/launches_timed.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("$OUT/launches_timed.csv")))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:30], float(r[vi].replace(',',''))/1e3) for r in rows[start:] if len(r)>vi]
for k,v in seq[:18]: print(f"{k:32s} {v:8.1f} us")
PY
