#!/bin/bash
# ncu launch lists (gpu__time_duration) of the multi-CTA path at B = 64, K = 1e6, both modes
mkdir -p gpurun_out
for mode in fast exact; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2q_large_${mode}_B64_K1e6_launches.csv python scripts/profile_step.py --mode $mode --batch 64 --particles 1000000 --launches 3 > gpurun_out/r2q_${mode}.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r2q_large_${mode}_B64_K1e6_launches.csv")) if len(r)>10]
hdr=rows[0]; H={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[H["Kernel Name"]][:60]; m=r[H["Metric Name"]]; v=float(r[H["Metric Value"]].replace(",",""))
    agg.setdefault(k,collections.defaultdict(list))[m].append(v)
print("$mode")
for k,d in agg.items():
    t=d.get("gpu__time_duration.sum",[0]); rd=d.get("dram__bytes_read.sum",[0]); wr=d.get("dram__bytes_write.sum",[0])
    print("  %-60s n=%d  us %.1f  dram MB rd %.0f wr %.0f"%(k,len(t),sum(t)/len(t)/1e3 if max(t)>1e3 else sum(t)/len(t), sum(rd)/len(rd)/ (1e6 if max(rd)>1e6 else 1), sum(wr)/len(wr)/(1e6 if max(wr)>1e6 else 1)))
PY
done
