#!/bin/bash
# A/B visit: parity tests under the new default, then bench.py per value of one SGB_* switch.   bash scripts/gpu_ab.sh TAG SWITCH v1 v2 ...
TAG=$1; SW=$2; shift 2
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_fused_allreduce.py -q -m gpu -x > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log; echo "t=$(( $(date +%s) - T0 ))s"
for v in "$@"; do
  env $SW=$v timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > $OUT/bench_${SW}_$v.json 2> $OUT/bench_${SW}_$v.err; echo "rc=$?" >> $OUT/bench_${SW}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_${SW}_$v.json")); print("$SW=$v value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "warm", round(d["value_l2_warm"],1), "per_pose", [round(x,4) for x in d["per_pose_ms"]], "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("fail", e); print(open("$OUT/bench_${SW}_$v.err").read()[-1500:])
PY
done
echo "t=$(( $(date +%s) - T0 ))s"
