"""Top stalled SASS instructions of one kernel in an ncu report (development aid).
usage: ncu_hot.py report.ncu-rep kernel-regex [N]"""
import csv, io, subprocess, sys
rep, rx = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# several kernels may follow each other: split at header rows
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["data"].append(r)
for b in blocks[:1]:
    h, d = b["hdr"], b["data"]
    iS, iN, iE = h.index("Source"), h.index("Warp Stall Sampling (Not-issued Samples)"), h.index("Instructions Executed")
    num = lambda x: int(float(x)) if x not in ("", None) else 0
    tot = sum(num(r[iN]) for r in d)
    print(b["name"][:80], "instructions", len(d), "not-issued samples", tot)
    top = sorted(range(len(d)), key=lambda i: -num(d[i][iN]))[:N]
    for i in sorted(top):
        print(f"{i:5d} {num(d[i][iN]):7d} {100.0*num(d[i][iN])/max(tot,1):5.1f}% exec {num(d[i][iE]):9d}  {d[i][iS][:100]}")
