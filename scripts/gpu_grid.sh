#!/bin/bash
# Grid front-end / factor kernel tuning round: parity tests, A/B switches, per-launch times and one full ncu capture.
#   usage: gpu_grid.sh TAG "ENV1=a ENV2=b" "ENV1=c" ...   (each quoted argument = one bench variant)
TAG=${1:-r01p}
shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1200 python -m pytest tests -q -m gpu -x >> $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
show() {
python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2 | value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "warm", round(d["value_l2_warm"],1), "launches", d["gpu_launches"], "parity", d.get("parity_rel_H"))
except Exception as e:
    print("fail", e); print(open("$1".replace(".json",".err")).read()[-1500:])
PY
}
k=0
for v in "$@"; do
k=$((k+1))
env $v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_v$k.json 2> $OUT/bench_v$k.err
show $OUT/bench_v$k.json "v$k [$v]"
echo "-- launch list [$v]"
env $v timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"packet_search|factor_reduce|grid_probe|pending_search" -s 40 -c 20 --csv --log-file $OUT/launches_v$k.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("$OUT/launches_v$k.csv")))
for i,r in enumerate(rows):
    if 'Kernel Name' in r: hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
seq=[(r[ki][:24], float(r[vi].replace(',',''))/1e3) for r in rows[start:] if len(r)>vi]
print("  ".join(f"{k.split('(')[0][:14]}={v:.1f}" for k,v in seq[:20]))
PY
done
echo "== ncu full (last variant)"
env $v timeout 900 ncu --set full --import-source on --clock-control none -k regex:"grid_probe|pending_search|packet_search|factor_reduce" -s 40 -c 8 -o $OUT/full -f \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/ncu_full.log 2>&1
ls $OUT | head -40
