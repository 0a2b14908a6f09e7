"""Per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep captured with --import-source on
(read here, no GPU needed).   python tools/ncu_lines.py gpurun_out/prof.ncu-rep [min_pct] [out.csv]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, x in enumerate(rows) if x and x[0] == "Line No"][0]
h = rows[hi]
ii, sa = h.index("Instructions Executed"), h.index("# Samples")
per = {}
order = []
tot_i = tot_s = 0
cur = None
for row in rows[hi + 1:]:
    if len(row) < len(h):
        continue
    if row[0].strip():
        cur = (row[0], row[1])
        if cur not in per:
            per[cur] = [0, 0]
            order.append(cur)
    try:
        n, s = int(row[ii]), int(row[sa])
    except ValueError:
        continue
    if not row[2].strip():
        continue  # the cuda line's own summary row (would double count)
    per[cur][0] += n
    per[cur][1] += s
    tot_i += n
    tot_s += s
out = [f"# {rep}: {tot_i} warp-instructions, {tot_s} stall samples", "line,inst_pct,sample_pct,source"]
for k in order:
    n, s = per[k]
    if tot_i and 100.0 * n / tot_i >= min_pct or tot_s and 100.0 * s / tot_s >= min_pct:
        out.append(f"{k[0]},{100.0 * n / tot_i:.2f},{100.0 * s / max(tot_s, 1):.2f},\"{k[1].strip()[:140]}\"")
txt = "\n".join(out)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(txt + "\n")
print(txt)
