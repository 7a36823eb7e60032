"""DRAM traffic per launch of K-fwd / K-inv from the ncu csv of tools/gpu_job_profiles.sh -> profiles/r2_final_traffic.json.

    python tools/traffic_json.py gpurun_out/r2c_traffic.csv > profiles/r2_final_traffic.json
"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")) if len(r) > 10]
hdr = rows[0]
ki, gi, mi, vi = (hdr.index(k) for k in ("Kernel Name", "Grid Size", "Metric Name", "Metric Value"))
acc = collections.defaultdict(lambda: collections.defaultdict(list))
images = None
for r in rows[1:]:
    name = "k_forward" if "k_forward" in r[ki] else "k_inverse"
    acc[name][r[mi]].append(float(r[vi].replace(",", "")))
    images = int(re.findall(r"\d+", r[gi])[-1])
out = {"config": {"images": images, "width": 1920, "height": 1080, "channels": 3, "quality": 50},
       "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "
              "regex:k_forward3|k_inverse4 python bench.py --steps 1 --warmup 1 (tools/gpu_job_profiles.sh); averages per launch",
       "algorithmic_bytes_per_launch": 2 * 3 * 1920 * 1080 * images}
for name, m in acc.items():
    rd, wr = (sum(m[k]) / len(m[k]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    out[name] = {"traffic": rd + wr, "read": rd, "write": wr, "ms_under_ncu": sum(m["gpu__time_duration.sum"]) / len(m["gpu__time_duration.sum"]) / 1e6,
                 "launches": len(m["dram__bytes_read.sum"])}
print(json.dumps(out, indent=1))
