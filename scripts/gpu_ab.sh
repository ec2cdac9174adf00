#!/bin/bash
# A/B of profiling-library switches on the bench line (no extras, no CPU leg).  Usage: bash scripts/gpu_ab2.sh <tag> "name:ENV=v,ENV=v" ...
# Each variant runs in its own process with SGB_LIBRARY=prof (libsgicp_b200_prof.so reads the SGB_* switches); "product" = the shipped library.
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  [ "$envs" = "$spec" ] && envs=""
  (
    if [ "${name#prev}" != "$name" ]; then export SGB_LIBRARY=prev; elif [ "${name#product}" = "$name" ]; then export SGB_LIBRARY=prof; fi
    IFS=',' read -ra kv <<< "$envs"
    for e in "${kv[@]}"; do [ -n "$e" ] && export "$e"; done
    timeout 300 python bench.py --steps 40 --warmup 5 --no-extras --no-cpu-baseline > $OUT/$name.json 2> $OUT/$name.err
  )
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/$name.json").read().strip().splitlines()[-1])
    print("%-28s ms/step %.4f  per pose %s  e2e %.3f  err %.4f" % ("$name", d["ms_per_step"], " ".join("%.4f"%x for x in d["per_pose_ms"]), d["e2e"]["ms_per_step"], d["error_ms"]["value"]))
except Exception as e:
    print("$name", "FAILED", e); print(open("$OUT/$name.err").read()[-1500:])
PY
done
