#!/bin/bash
# bench line only (ours [+ reference arm]).  Usage: bash scripts/gpu_bench.sh <tag> [N gpus] [extra bench args]
TAG=${1:-r02b}; N=${2:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 20 --warmup 3 ${@:3} > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 ${@:3} > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err; echo "rc=$?" >> $OUT/bench_n$N.err
fi
echo "t=$(( $(date +%s) - T0 ))s"
tail -c 1500 $OUT/bench*.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, {k:d.get(k) for k in ("value","ms_per_step","per_pose_ms","allreduce_check_rel","comm_wait_us")})
    print(" e2e", d["e2e"]["ms_per_step"], "error_ms", d["error_ms"]["value"], "roofline", d["roofline"]["frac"], "pose_vs_ref", d.get("pose_error_vs_reference"))
    for k in ("c3","c4","c5"):
        print(" ",k, json.dumps(d.get(k))[:1800])
PY
