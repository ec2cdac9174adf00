#!/bin/bash
# A/B visit: parity tests (default switches, then under the first non-default spec), then bench.py once per spec.
#   bash scripts/gpu_ab.sh TAG SPEC [SPEC ...]     SPEC = "-" (defaults) or "VAR=val,VAR2=val2"
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_fused_allreduce.py tests/test_host_mirror.py -q -m gpu -x > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log; echo "t=$(( $(date +%s) - T0 ))s"
LAST="${@: -1}"
if [ "$LAST" != "-" ]; then
  env $(echo $LAST | tr ',' ' ') timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "structures or far_and or synthetic_200k or golden or degenerate" > $OUT/pytest_variant.log 2>&1; echo "rc=$?" >> $OUT/pytest_variant.log
  echo "variant [$LAST]:"; tail -3 $OUT/pytest_variant.log; echo "t=$(( $(date +%s) - T0 ))s"
fi
i=0
for spec in "$@"; do
  i=$((i+1))
  envs=""; [ "$spec" != "-" ] && envs=$(echo $spec | tr ',' ' ')
  env $envs timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > $OUT/bench_$i.json 2> $OUT/bench_$i.err; echo "rc=$?" >> $OUT/bench_$i.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$i.json")); print("[$spec] value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "warm", round(d["value_l2_warm"],1), "per_pose", [round(x,4) for x in d["per_pose_ms"]], "clocks", d["clocks"]["sm_mhz"])
except Exception as e:
    print("[$spec] fail", e); print(open("$OUT/bench_$i.err").read()[-1500:])
PY
done
echo "t=$(( $(date +%s) - T0 ))s"
