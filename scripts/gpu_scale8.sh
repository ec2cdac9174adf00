#!/bin/bash
# One visit to an 8-GPU box: the fused all-reduce path of bench.py at N = 8 (and 4), as the round-end scaling run launches it.
TAG=${1:-r01_scale}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 3 \
    > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1])
    print("N=$N value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), d["config"].get("allreduce"), d["pose_error_vs_gt"])
except Exception as e:
    print("fail", e); print(open("$OUT/bench_n$N.err").read()[-2500:])
PY
done
