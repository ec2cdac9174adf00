#!/bin/bash
TAG=${1:-r01h}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1200 python -m pytest tests -q -m gpu >> $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
show() {
python - <<PY
import json
try:
    d=json.load(open("$1"))
    print("$2 value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "warm", round(d["value_l2_warm"],1), "e2e", round(d["e2e"]["value"],1), "launches", d["gpu_launches"], "iters", d["pose_error_vs_gt"]["gn_iterations"], d["config"]["covariances"][:20])
except Exception as e:
    print("fail", e); print(open("$1".replace(".json",".err")).read()[-1500:])
PY
}
for curve in 1 0; do
  name=curve${curve}
  SGB_CURVE=$curve timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --covs analytic > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  show $OUT/bench_$name.json "curve=$curve (1=hilbert) analytic covs"
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_knn.json 2> $OUT/bench_knn.err
show $OUT/bench_knn.json "default (hilbert, knn covs)"
