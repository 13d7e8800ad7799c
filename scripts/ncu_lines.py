#!/usr/bin/env python
"""Per-source-line instruction / stall-sample breakdown of an .ncu-rep (needs -lineinfo + --import-source on).
    python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top_n]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; agg = collections.defaultdict(lambda: [0, 0, ""]); tot = 0; S = 0
for r in csv.reader(io.StringIO(raw)):
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Line No": hdr = r; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if hdr and len(r) == len(hdr):
        try: e = int(r[iE]); s = int(r[iS])
        except ValueError: continue
        key = (cur, r[0]); agg[key][0] += e; agg[key][1] += s; agg[key][2] = r[1][:100]; tot += e; S += s
print("instructions", tot, "samples", S)
byfile = collections.Counter()
for (f, ln), (e, s, src) in agg.items(): byfile[f] += e
print({k: f"{v / tot * 100:.1f}%" for k, v in byfile.items()})
for (f, ln), (e, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{e / tot * 100:5.1f}% inst {s / max(S,1) * 100:5.1f}% samp  {f}:{ln}  {src.strip()}")
