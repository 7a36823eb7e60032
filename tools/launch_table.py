"""Per-step kernel table from an ncu launch list of bench.py (profiles/r1_final_launches.txt)."""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("himgcu::", "").replace("void ", "")
    agg.setdefault((name, r[gi]), []).append(float(r[vi].replace(",", "")) / 1e3)
# the 512-image step launches every (kernel, grid) pair below exactly once; the 5-image sub-batches of
# the host-buffer leg and other small launches are left out
step = [(k, sum(v) / len(v), len(v)) for k, v in agg.items() if len(v) in (8, 16)]
tot = sum(a for _, a, _ in step)
print("# One encode+decode step of bench.py (512 x 1920x1080x3, q50) under")
print("#   ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:k_(forward|inverse|huff|dec|lowres|lres)")
print("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's event timings, not absolutes")
print("# average duration of every (kernel, grid) launched once per step; total %.1f us" % tot)
for k, a, n in sorted(step, key=lambda x: -x[1]):
    print(f"{k[0]:28s} grid={k[1]:20s} avg={a:10.1f} us  share={a / tot:6.3f}  (n={n} launches captured)")
