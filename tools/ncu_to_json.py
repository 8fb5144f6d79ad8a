"""Per-launch figures of one kernel from an ncu report -> small JSON under profiles/ (read by bench.py).
usage: ncu_to_json.py report.ncu-rep 'kernel-name-substring' frames out.json [launch_index]"""
import csv, io, json, subprocess, sys
rep, pat, frames, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
which = int(sys.argv[5]) if len(sys.argv) > 5 else 0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {k: i for i, k in enumerate(hdr)}
cands = [r for r in rows[2:] if pat in r[col["Kernel Name"]]]
r = cands[which]


def num(key, scale=1.0):
    v = r[col[key]].replace(",", "")
    try:
        return float(v) * scale
    except ValueError:
        return None


unit_scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}
dram = sum(num(k, unit_scale.get(units[col[k]], 1.0)) or 0.0 for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
t = num("gpu__time_duration.sum")
tu = units[col["gpu__time_duration.sum"]]
t_ms = t * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(tu, 1.0)
inst = num("smsp__inst_executed.sum")
ratio = num("smsp__thread_inst_executed_per_inst_executed.ratio")
res = {"kernel": r[col["Kernel Name"]], "frames": frames, "report": rep.split("/")[-1],
       "how": "ncu --set full --clock-control none (cold-cache, serialised launch; per-launch values)",
       "ncu_launch_ms": t_ms, "dram_bytes": dram,
       "thread_inst": inst * ratio if inst and ratio else None,
       "registers_per_thread": num("launch__registers_per_thread"),
       "shared_mem_per_block_kb": num("launch__shared_mem_per_block_dynamic"),
       "blocks_per_sm_limit_smem": num("launch__occupancy_limit_shared_mem"),
       "blocks_per_sm_limit_regs": num("launch__occupancy_limit_registers"),
       "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active"),
       "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "fma_pipe_pct": num("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
       "alu_pipe_pct": num("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
       "xu_pipe_pct": num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
       "lsu_pipe_pct": num("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
       "smem_wavefronts": num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
       "smem_bank_conflicts": num("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum")}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps(res))
