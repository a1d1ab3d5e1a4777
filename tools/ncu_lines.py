"""Per-source-line instruction counts / stall samples from an .ncu-rep (needs -lineinfo + --import-source on).

    python tools/ncu_lines.py rep.ncu-rep [--kernel regex] [--top 40]
"""
import argparse
import csv
import io
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("--kernel", default="")
ap.add_argument("--top", type=int, default=40)
a = ap.parse_args()
cmd = ["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if a.kernel:
    cmd += ["--kernel-name", "regex:" + a.kernel]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
fname, func, hdr, lines, seen = None, None, None, [], set()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0] not in ("", "-") and r[0].isdigit():
        ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        key = (func, fname, int(r[0]))
        if key in seen:
            continue
        seen.add(key)
        try:
            lines.append((int(r[ci]), int(r[si]), fname, int(r[0]), r[1].strip()[:110], func))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1
tots = sum(l[1] for l in lines) or 1
print(f"total warp-instructions {tot}, samples {tots}")
for c, s, f, ln, src, fn in sorted(lines, reverse=True)[: a.top]:
    print(f"{c / tot * 100:5.1f}% inst {s / tots * 100:5.1f}% stall  {f}:{ln:<4d} {src}")
