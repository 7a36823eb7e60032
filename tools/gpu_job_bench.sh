#!/bin/bash
# One GPU job: full GPU test suite + the default bench run; results under gpurun_out/<tag>_*.
tag=${1:-job}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.txt 2>&1
tail -3 gpurun_out/${tag}_pytest.txt
( time python bench.py --steps 3 ) > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_bench.json") if l.startswith("{")][-1])
except Exception as e:
    print("no bench line:", e); raise SystemExit
for k in ("value", "ms_per_step", "encode_mps", "decode_mps", "parity_checked", "parity_images_checked_against_reference_hashes",
          "multi_gpu_parity", "gpu_launches", "kernel_time_shares", "clocks"):
    print(k, d.get(k))
print("e2e", {k: v for k, v in d["e2e"].items() if k != "api"})
print("roofline", d["roofline"])
print("cpu", d["cpu_baseline"])
for k in ("c2", "c3", "c5"):
    print(k, d.get(k))
PY
