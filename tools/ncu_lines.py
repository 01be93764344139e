"""Per-source-line stall samples and executed instructions of one kernel from an .ncu-rep captured with
--import-source on (reads `ncu --page source --csv --print-source cuda,sass`).
  python tools/ncu_lines.py report.ncu-rep kernel_name [top]"""
import csv
import subprocess
import sys

rep, kern, top = sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
ci = {name: h.index(name) for name in ("Line No", "# Samples", "Instructions Executed")}
src_col = 1
agg = {}
line, text = None, ""
for r in rows[hi + 1:]:
    if len(r) < len(h):
        continue
    if r[0].strip():
        if not r[0].strip().isdigit():
            continue
        line, text = int(r[0]), r[src_col]
    try:
        s, e = int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0)
    except ValueError:
        continue
    a = agg.setdefault(line, [0, 0, text])
    a[0] += s
    a[1] += e
ts, te = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
print(f"{kern}: {ts} samples, {te} warp instructions")
for ln, (s, e, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:5d} samples {s / max(ts, 1) * 100:5.1f}%  instr {e / max(te, 1) * 100:5.1f}%  {t.strip()[:110]}")
