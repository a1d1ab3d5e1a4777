"""Average the per-launch metrics of a multi-launch .ncu-rep (read here, no GPU) and print them as JSON:
    python tools/ncu_ring_json.py rep.ncu-rep [--kernel substr]"""
import argparse
import csv
import io
import json
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--kernel", default="")
a = ap.parse_args()
M = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
     "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slot_pct", "smsp__inst_executed.sum": "warp_instructions",
     "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
     "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
     "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "lsu_pipe_pct",
     "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
     "launch__registers_per_thread": "registers", "launch__grid_size": "grid", "launch__block_size": "block",
     "launch__occupancy_limit_registers": "ctas_per_sm_by_registers", "lts__t_bytes.sum": "l2_bytes",
     "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts"}
raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
acc, n, name = {}, 0, None
def num(s):
    return float(s.replace(",", ""))
scale = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}
for r in rows[2:]:
    kn = r[hdr.index("Kernel Name")]
    if a.kernel and a.kernel not in kn:
        continue
    name = kn
    n += 1
    for m, out in M.items():
        if m in hdr:
            i = hdr.index(m)
            acc[out] = acc.get(out, 0.0) + num(r[i]) * scale.get(units[i], 1.0)
st = {}
for s in ["barrier", "short_scoreboard", "long_scoreboard", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "branch_resolving", "not_selected", "no_instruction", "dispatch_stall", "membar", "sleeping", "drain"]:
    m = f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio"
    if m in hdr:
        st[s] = round(sum(num(r[hdr.index(m)]) for r in rows[2:] if not a.kernel or a.kernel in r[hdr.index("Kernel Name")]) / max(n, 1), 3)
out = {k: round(v / n, 3) for k, v in acc.items()}
out["dram_bytes"] = round(out.get("dram_read_bytes", 0) + out.get("dram_write_bytes", 0), 1)
out["issue_slot_frac"] = round(out.get("issue_slot_pct", 0) / 100, 4)
out["launches_averaged"] = n
out["kernel"] = name
out["stalls_warps_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1]))
print(json.dumps(out, indent=1))
