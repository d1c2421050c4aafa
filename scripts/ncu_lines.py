"""Summarise an ncu report per CUDA source line: stall samples and executed instructions.

    python scripts/ncu_lines.py gpurun_out/step_exact.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        try:
            samples = int(d.get("# Samples", "0") or 0)
            inst = int(d.get("Instructions Executed", "0") or 0)
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        a = agg.setdefault(key, [0, 0, r[1].strip()[:110]])
        a[0] += samples
        a[1] += inst
tot_s = sum(a[0] for a in agg.values()) or 1
tot_i = sum(a[1] for a in agg.values()) or 1
print("total samples %d, total warp-instructions %d" % (tot_s, tot_i))
for (f, line), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * s / tot_s, 100.0 * i / tot_i, f, line, src))
