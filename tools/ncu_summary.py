"""Summarise an .ncu-rep (read here, no GPU): per-kernel headline metrics + instruction mix + hot SASS.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--hot 60] [--kernel regex]
"""
import argparse
import collections
import csv
import io
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--hot", type=int, default=0, help="print SASS instructions executed more than this many times per warp")
ap.add_argument("--kernel", default="")
a = ap.parse_args()

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
STALL = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALLS = ["barrier", "short_scoreboard", "long_scoreboard", "wait", "math_pipe_throttle", "mio_throttle", "lg_throttle",
          "branch_resolving", "not_selected", "no_instruction", "dispatch_stall", "membar", "sleeping", "drain"]

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    if a.kernel and a.kernel not in name:
        continue
    print("==", name[:110])
    for w in WANT:
        if w in hdr:
            print(f"   {w:75s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
    st = [(s, float(r[hdr.index(STALL % s)])) for s in STALLS if (STALL % s) in hdr]
    print("   stalls (warps per issue):", ", ".join(f"{s}={v:.2f}" for s, v in sorted(st, key=lambda x: -x[1]) if v > 0.05))

src = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + a.kernel] if a.kernel else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# the CSV concatenates kernels: "Kernel Name" row, header row, then instructions
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None:
        cur["rows"].append(r)
seen = set()
for b in blocks:
    if b["name"] in seen:
        continue
    seen.add(b["name"])
    h = b["hdr"]
    ci, si, sm = h.index("Instructions Executed"), h.index("Source"), h.index("# Samples")
    data = []
    for r in b["rows"]:
        try:
            data.append((r[si], int(r[ci]), int(r[sm])))
        except (ValueError, IndexError):
            pass
    tot, tots = sum(d[1] for d in data), max(1, sum(d[2] for d in data))
    ops, opsmp = collections.Counter(), collections.Counter()
    for s, c, m in data:
        t = s.split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        ops[op] += c
        opsmp[op] += m
    print("== SASS", b["name"][:100])
    print("   static instructions", len(data), " executed (warp-level)", tot)
    print("   mix:", ", ".join(f"{o} {c / tot * 100:.1f}% (stall {opsmp[o] / tots * 100:.0f}%)" for o, c in ops.most_common(16)))
    if a.hot:
        mx = max(d[1] for d in data)
        print(f"   hot instructions (> {a.hot}x the least-executed non-zero count):")
        base = min(d[1] for d in data if d[1] > 0)
        for s, c, m in data:
            if c > a.hot * base:
                print(f"     {c / base:9.1f}x  stall {m / tots * 100:4.1f}%  {s[:100]}")
