#!/bin/bash
# Experiment visit: new degenerate-density test, bench (clock sampling), full ncu capture of the packet search at the two misaligned poses.
TAG=${1:-r01ai}
OUT=gpurun_out/$TAG
mkdir -p $OUT
T0=$(date +%s)
timeout 200 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "degenerate or structures_agree or far_and" > $OUT/pytest.log 2>&1; echo "rc=$?" >> $OUT/pytest.log
tail -6 $OUT/pytest.log; echo "t=$(( $(date +%s) - T0 ))s"
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "rc=$?" >> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json")); print("value", d["value"], "clocks", d["clocks"], "per_pose", d["per_pose_ms"])
PY
echo "t=$(( $(date +%s) - T0 ))s"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:packet_search -s 0 -c 2 -f -o $OUT/prof_packet \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_packet.log 2>&1
echo "rc=$?"; ls -la $OUT; echo "t=$(( $(date +%s) - T0 ))s"
