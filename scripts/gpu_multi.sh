#!/bin/bash
# N-GPU validation of bench.py under torchrun (weak scaling: 1M source points per GPU): the all-reduce of H|b|e fused into
# the reduction kernel (peer mailboxes over NVLink) vs the NCCL all_reduce of the same 44 doubles, plus the reference arm.
N=${1:-2}
TAG=${2:-r01_multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
show() {
python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1])
    print("$2 | value", round(d["value"],1), "ms", round(d["ms_per_step"],4), "kernel_ms", round(d["roofline"]["kernel_ms"],4), "e2e", round(d["e2e"]["value"],1), "allreduce:", d["config"].get("allreduce"), "pose err", d["pose_error_vs_gt"])
except Exception as e:
    print("fail", e); print(open("$1".replace(".json",".err")).read()[-2500:])
PY
}
echo "== fused test on the GPUs of this box (single process)"
timeout 600 python -m pytest tests/test_gpu_fused_allreduce.py -q -x > $OUT/pytest_fused.log 2>&1; echo "rc=$?"; tail -3 $OUT/pytest_fused.log
for mode in 1 0; do
SGB_FUSED_ALLREDUCE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$mode bench.py --gpus $N --steps 20 --warmup 3 \
    > $OUT/bench_n${N}_fused$mode.json 2> $OUT/bench_n${N}_fused$mode.err
echo "rc=$?"
show $OUT/bench_n${N}_fused$mode.json "N=$N fused=$mode"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err
echo "rc=$?"; tail -c 400 $OUT/bench_ref_n$N.json
