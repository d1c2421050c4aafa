"""Top source lines of an ncu report by executed warp instructions (cumulative %).
    python scripts/ncu_inst.py report.ncu-rep [n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No": hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        try: s = int(d.get("# Samples") or 0); i = int(d.get("Instructions Executed") or 0)
        except ValueError: continue
        a = agg.setdefault((cur, int(r[0])), [0, 0, r[1].strip()[:100]]); a[0] += s; a[1] += i
ti = sum(a[1] for a in agg.values()) or 1; ts = sum(a[0] for a in agg.values()) or 1
print("total warp-inst %d" % ti)
cum = 0
for (f, line), (s, i, src) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    cum += i
    print("%5.1f%% inst (cum %5.1f%%) %5.1f%% smp  %s:%d  %s" % (100.0 * i / ti, 100.0 * cum / ti, 100.0 * s / ts, f, line, src))
