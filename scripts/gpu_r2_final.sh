#!/bin/bash
# Round-2 closing run on one B200: what the driver does (GPU tests, smoke, bench, reference arm), then the evidence for
# profiles/: ncu launch list of bench.py, ncu --set full of the exact row kernel, clock-stamp timeline, compute-sanitizer.
# Needs build/variants/libaesmc_{timeline,sync2}.so (scripts/build_variants.sh timeline "-DAESMC_X_TIMELINE=1" sync2 "-DAESMC_X_CHAIN_SYNC=2").
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r2f_pytest.log
tail -3 gpurun_out/r2f_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2f_bench.json").read().strip().splitlines()[-1])
print("value %.4e ms %.2f frac %s e2e %.4e launches %s clocks %s"%(d["value"],d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["gpu_launches"],d.get("clocks")))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r2f_bench_reference.json
cut -c1-300 gpurun_out/r2f_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f_bench_launches.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2f_ncu_bench.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/r2f_bench_launches.csv")) if len(r)>10]
hdr=rows[0]; H={h:i for i,h in enumerate(hdr)}
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[H["Kernel Name"]][:70]; v=float(r[H["Metric Value"]].replace(",",""))
    u=r[H["Metric Unit"]]
    v = v/1e3 if u.startswith("ns") else v
    agg.setdefault(k,[]).append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:8]:
    print("  %-70s n=%d  avg us %.1f  share %.1f%%"%(k,len(v),sum(v)/len(v),100*sum(v)/tot))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:smc_step_x -s 4 -c 1 -f -o gpurun_out/r2f_x python scripts/profile_step.py --mode exact > gpurun_out/r2f_ncu.log 2>&1
tail -1 gpurun_out/r2f_ncu.log
AESMC_B200_LIB=$PWD/build/variants/libaesmc_timeline.so timeout 200 python scripts/timeline_step.py > gpurun_out/r2_timeline.txt 2>&1
head -3 gpurun_out/r2_timeline.txt
timeout 200 python scripts/bench_step_variant.py --label final 2>/dev/null > gpurun_out/r2f_step.json; cut -c1-330 gpurun_out/r2f_step.json
bash scripts/gpu_r2_t.sh > gpurun_out/r2f_sanitizer.log 2>&1; tail -22 gpurun_out/r2_sanitizer.txt
