#!/bin/bash
TAG=${1:-r01j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
echo "== pytest -m gpu" | tee $OUT/pytest.log
timeout 1200 python -m pytest tests -q -m gpu >> $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -6 $OUT/pytest.log
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
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_devkd.json 2> $OUT/bench_devkd.err
show $OUT/bench_devkd.json "device kd (refined)"
SGB_TREE=lbvh timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench_lbvh.json 2> $OUT/bench_lbvh.err
show $OUT/bench_lbvh.json "device LBVH"
python - <<'PY'
# build-time comparison of the three constructions on the 1M target
import time, numpy as np, os
import small_gicp_b200 as sg
from small_gicp_b200.synthetic import make_pair
tgt, src, T = make_pair(1_000_000)
for name, env in (("device-kd", None), ("lbvh", "lbvh"), ("host-kd", "host")):
    if env: os.environ["SGB_TREE"] = env
    else: os.environ.pop("SGB_TREE", None)
    ctx = sg.Context(0); ctx.set_target(tgt); ctx.build_target_kdtree(); ctx.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); ctx.build_target_kdtree(); ctx.synchronize(); ts.append(time.perf_counter() - t0)
    print(name, "build ms", [round(t * 1e3, 2) for t in ts]); ctx.close()
PY
