#!/usr/bin/env python
"""Top stalled SASS instructions of each distinct kernel in an .ncu-rep (needs ncu on PATH)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 14
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
done = set()
for s in secs:
    if s["name"] in done or not s["rows"]: continue
    done.add(s["name"])
    h = s["rows"][0]; ix = {n: i for i, n in enumerate(h)}
    data = []
    for r in s["rows"][1:]:
        try: data.append((int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]]), r[ix["Source"]], r))
        except (ValueError, IndexError, KeyError): pass
    tot = sum(d[0] for d in data) or 1
    print("====", s["name"][:80], "samples", tot)
    for d in sorted(data, key=lambda d: -d[0])[:topn]:
        r = d[3]
        st = {n: r[ix[n]] for n in h if n.startswith("stall_") and "Not Issued" not in n and r[ix[n]] not in ("0", "")}
        t2 = sorted(st.items(), key=lambda kv: -float(kv[1]))[:2]
        print(f"  {100*d[0]/tot:5.1f}% exec={d[1]:9d} {d[2][:64]:64s} {t2}")
