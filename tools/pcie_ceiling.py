"""Host <-> device copy ceiling of the box, all GPUs at once: what bounds the e2e leg of bench.py.

    python tools/pcie_ceiling.py                                   (1 GPU)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_ceiling.py

Every rank copies page-locked buffers to and from ITS GPU with plain cudaMemcpyAsync -- no coding -- host to
device only, device to host only, and both directions at once; all ranks start together (barrier) and the
slowest rank ends the measurement.  Prints one JSON line (rank 0) with GB/s per direction summed over the GPUs,
plus the topology facts that explain it (NUMA node of every GPU, CPU affinity, memory policy).

--numa-local  binds the page-locked allocation to the NUMA node of the rank's GPU (set_mempolicy) before it is
              made: on a two-socket host the default first-touch placement puts every buffer on the node the
              process happens to run on, and the copies of the far GPUs cross the socket link."""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import torch
import torch.distributed as dist


def gpu_numa_node(index):
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True).stdout.strip().lower()
        bdf = bdf[-12:] if len(bdf) > 12 else bdf  # 00000000:1B:00.0 -> 0000:1b:00.0
        return int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
    except (OSError, ValueError):
        return -1


def bind_memory_to_node(node):
    """set_mempolicy(MPOL_BIND, {node}) for this thread (x86-64 syscall 238); False if it is not allowed."""
    if node < 0:
        return False
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(1 << node)
    rc = libc.syscall(238, 2, ctypes.byref(mask), ctypes.c_ulong(65))
    return rc == 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=8)
    ap.add_argument("--numa-local", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    node = gpu_numa_node(local)
    bound = bind_memory_to_node(node) if args.numa_local else False
    n = args.mb << 20
    h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)  # first touch
    h_out.fill_(2)
    d_in = torch.empty(n, dtype=torch.uint8, device=dev)
    d_out = torch.ones(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(h2d, d2h):
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        barrier()
        return time.perf_counter() - t0

    res = {}
    for name, a, b in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
        run(a, b)
        t = run(a, b)
        tt = torch.tensor([t], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        res[name + "_gbs_per_direction_all_gpus"] = world * args.reps * n / float(tt.item()) / 1e9
    info = {"gpu": local, "numa_node": node, "bound": bound,
            "cpus": len(os.sched_getaffinity(0)), "cpu_list": sorted(os.sched_getaffinity(0))[:4] + ["..."] + sorted(os.sched_getaffinity(0))[-2:]}
    infos = [None] * world
    if world > 1:
        dist.all_gather_object(infos, info)
    else:
        infos = [info]
    if rank == 0:
        nodes = []
        try:
            nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
        except OSError:
            pass
        out = {"n_gpus": world, "buffer_mb": args.mb, "reps": args.reps, "numa_local_requested": args.numa_local,
               "host_numa_nodes": nodes, "ranks": infos}
        out.update(res)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
