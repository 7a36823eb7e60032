#!/bin/bash
# 8-GPU job: topology, copy ceiling on 1/2/4/8 GPUs (default and NUMA-local page-locked memory), bench at N GPUs.
tag=${1:-multi}
mkdir -p gpurun_out
{
  nvidia-smi topo -m
  echo "--- lscpu"; lscpu | head -25
  echo "--- numa"; cat /sys/devices/system/node/online 2>/dev/null; numactl -H 2>/dev/null | head -20
  echo "--- affinity"; python -c "import os; print(len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0)))"
  echo "--- mems"; grep -i "mems_allowed_list\|cpus_allowed_list" /proc/self/status
  free -g | head -3
} > gpurun_out/${tag}_topo.txt 2>&1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) "${@:2}"; }
for n in 1 2 4 8; do
  run $n tools/pcie_ceiling.py 2>/dev/null | grep '^{' > gpurun_out/${tag}_ceiling_${n}.json
  run $n tools/pcie_ceiling.py --numa-local 2>/dev/null | grep '^{' > gpurun_out/${tag}_ceiling_numa_${n}.json
done
cat gpurun_out/${tag}_ceiling_*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['n_gpus'], 'numa_local' if d['numa_local_requested'] else 'default', {k: round(v, 1) for k, v in d.items() if k.endswith('all_gpus')}, [ (r['numa_node'], r['bound']) for r in d['ranks']])
"
for n in 8 2; do
  ( time run $n bench.py --gpus $n --steps 3 ) > gpurun_out/${tag}_bench_${n}.json 2> gpurun_out/${tag}_bench_${n}.err
  tail -4 gpurun_out/${tag}_bench_${n}.err
  python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_bench_${n}.json") if l.startswith("{")][-1])
    print("N=${n}", "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "serial", round(d["e2e"]["serial_value"]), "parity", d["parity_images_checked_against_reference_hashes"], d["multi_gpu_parity"], "frac", round(d["roofline"]["frac"], 3), round(d["roofline"]["k_inverse"]["frac"], 3))
except Exception as e:
    print("no bench line", e)
PY
done
