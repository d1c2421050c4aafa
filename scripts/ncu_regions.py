"""Aggregate executed warp-instructions and stall samples of an ncu report by source file and line range.
    python scripts/ncu_regions.py report.ncu-rep
"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; per = collections.OrderedDict()
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        try: s = int(d.get("# Samples") or 0); i = int(d.get("Instructions Executed") or 0)
        except ValueError: continue
        k = (cur, int(r[0])); a = per.setdefault(k, [0, 0]); a[0] += s; a[1] += i
ti = sum(v[1] for v in per.values()) or 1; ts = sum(v[0] for v in per.values()) or 1
regions = {
 "common.cuh": [(1, 60, "misc"), (61, 83, "block_allreduce"), (84, 116, "np_expf"), (117, 141, "np_logf"), (142, 225, "fd_log1pf"), (226, 245, "count_below fp64"), (246, 270, "count_below filtered")],
}
agg = collections.OrderedDict()
for (f, line), (s, i) in per.items():
    name = None
    for lo, hi, n in regions.get(f, []):
        if lo <= line <= hi: name = f + ":" + n
    if name is None: name = f + ":" + str(line // 20 * 20) + "-" + str(line // 20 * 20 + 19)
    a = agg.setdefault(name, [0, 0]); a[0] += s; a[1] += i
print("total warp-inst %d, samples %d" % (ti, ts))
for n, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    if i * 200 > ti or s * 200 > ts: print("%5.1f%% inst %5.1f%% smp  %s" % (100.0 * i / ti, 100.0 * s / ts, n))
