"""Summarise the warp-stall samples of one kernel per CUDA source line, from
`ncu -i <rep> --page source --csv --print-source cuda,sass` (needs -lineinfo and --import-source on).
usage: python tools/ncu_stalls.py <rep> [top_n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rd = csv.reader(out.split("\n"))
cur_file, header = None, None
agg = collections.OrderedDict()
tot = 0
for r in rd:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        header = r; continue
    if header is None or len(r) < len(header):
        continue
    d = {}
    for k, v in zip(header, r):       # duplicate "Source" column: keep the first (CUDA text)
        d.setdefault(k, v)
    if not d["Line No"]:               # SASS row: counted under its CUDA line row already
        continue
    try:
        n = int(d["# Samples"])
    except ValueError:
        continue
    if n == 0:
        continue
    tot += n
    st = {k[6:]: int(d[k]) for k in header if k.startswith("stall_") and "Not Issued" not in k and d[k] not in ("", "0")}
    agg[(cur_file, int(d["Line No"]))] = (n, d["Source"].strip()[:90], sorted(st.items(), key=lambda kv: -kv[1])[:3])
print("total samples", tot)
for (f, ln), (n, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{n:6d} {100*n/tot:5.1f}%  {f}:{ln:<4d} {src}  {st}")
