#!/bin/bash
# DRAM bytes + duration of the three launches of sgb_linearize over one pass of the pose schedule (cheap ncu pass: three metrics).
# usage (under gpurun): bash scripts/gpu_traffic.sh <tag>
TAG=${1:-r02t}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SGB_BENCH_ROLL=5 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k 'regex:grid_probe|packet_search|factor_reduce' -s 45 -c 15 --csv --log-file $OUT/traffic.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > $OUT/traffic.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$OUT/traffic.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    name=r[4].split("(")[0].replace("void ","").replace("sgb::","")[:32]; metric=r[-3]; unit=r[-2]; val=float(r[-1].replace(",",""))
    agg.setdefault((r[0],name),{})[metric]=(val,unit)
tot=0.0
for (i,name),m in agg.items():
    rd=m.get("dram__bytes_read.sum",(0,""));wr=m.get("dram__bytes_write.sum",(0,""));du=m.get("gpu__time_duration.sum",(0,""))
    def mb(v,u): return v*{"byte":1e-6,"Kbyte":1e-3,"Mbyte":1,"Gbyte":1e3}.get(u,1)
    t=mb(*rd)+mb(*wr); tot+=t
    print("%-34s %8.2f %s  read %7.1f MB  write %6.1f MB"%(name,du[0],du[1],mb(*rd),mb(*wr)))
print("total DRAM MB over %d launches: %.1f -> per linearize %.1f MB"%(len(agg),tot,tot/(len(agg)/3.0)))
PY
