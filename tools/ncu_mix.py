"""Dynamic instruction mix of one kernel from an ncu report (development aid).
usage: ncu_mix.py report.ncu-rep kernel_regex units_per_launch [launch_index]"""
import csv, collections, re, subprocess, sys, io
rep, kre, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}",
                      "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][:2])
hdr = rows[1]; data = rows[2:]
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
ops = collections.Counter(); tot = 0
for r in data:
    try: e = int(r[iE])
    except Exception: continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS].strip())
    op = m.group(2) if m else r[iS][:10]
    ops[op] += e; tot += e
# the source page counts every instruction twice when the kernel was replayed; normalise by raw metric if given
print("total warp-instr (source page)", tot, " thread-instr per unit", tot * 32 / units)
for op, c in ops.most_common(28):
    print(f"{op:10s} {100*c/tot:6.2f}%   per unit {c*32/units:9.1f}")
