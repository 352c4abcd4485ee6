#!/usr/bin/env python
"""Hottest SASS lines (by warp-stall samples) of an .ncu-rep captured with --import-source on:
python scripts/ncu_hot.py file.ncu-rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
hdr = rows[0]
ia, isrc, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for n, r in enumerate(rows[1:]):
    try: s = int(r[ismp])
    except Exception: continue
    data.append((s, n, r))
tot = sum(d[0] for d in data)
print(f"total samples {tot}")
for s, n, r in sorted(data, reverse=True)[:top]:
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
    print(f"{100*s/tot:5.1f}%  #{n:5d}  {r[isrc][:90]:90s} {st}")
