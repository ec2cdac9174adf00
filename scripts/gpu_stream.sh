#!/bin/bash
# where a LiDAR frame's time goes: wall-clock split, then the ncu launch list summed by kernel name
OUT=gpurun_out/${1:-r02s}; mkdir -p $OUT
timeout 120 python scripts/stream_profile.py 3 2>&1 | tail -6
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/stream_launches.csv python scripts/stream_profile.py 2 > $OUT/stream.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$OUT/stream_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    name=r[ki].split("(")[0][-60:]
    tot[name]+=v; cnt[name]+=1
print("total launches", sum(cnt.values()), "total us", round(sum(tot.values())))
for n,v in tot.most_common(22): print("%9.1f us %5d x  %s" % (v, cnt[n], n))
PY
