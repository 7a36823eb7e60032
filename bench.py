#!/usr/bin/env python
"""bench.py -- HIMG encode+decode throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (config.workload): BASELINE.json configs[3] "batch of 4096 1920x1080 RGB images
encode+decode sharded across 1/2/4/8 B200", weak-scaled: every GPU codes `--images` (default 512)
1080p RGB images at quality 50, so 8 GPUs code the full 4096-image batch.  A "step" is one pass of
the hot path over the rank's batch: encode every image to a .himg bitstream, then decode every
bitstream back to pixels.

Printed JSON (one line, rank 0):
  value      megapixels (W*H, channels not counted) through encode AND decode per second, whole
             job, inputs resident in HBM, CUDA events, max over ranks
  e2e        the same through the host-buffer C-ABI calls (pinned host memory, H2D and D2H inside
             the timed region)
  roofline   K-fwd (k_forward: colour map + low-res subtract + WHT + quantise + map), algorithmic
             bytes = 2*nch per pixel, duration from CUDA events around every launch in the timed
             region, peak = MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the unmodified reference (oracle/_ref) on the host cores, bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W, H, NCH, QUALITY = 1920, 1080, 3, 50
WORKLOAD = "c4: batch of 1920x1080 RGB images, quality 50, YCbCr, encode+decode, sharded by image"


# ------------------------------------------------------------------------------------------------
# CPU side: the reference's own implementation on the host cores (oracle/_ref, else the port)
# ------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, seeds = args
    import numpy as np  # noqa: F401

    import oracle

    port = oracle.port()
    impl = oracle.ref() if kind == "reference" else port
    t_enc = t_dec = 0.0
    for seed in seeds:
        img = port.synth(W, H, NCH, seed, 6)
        t0 = time.perf_counter()
        packed = impl.encode(img, QUALITY, True)
        t1 = time.perf_counter()
        dec = impl.decode(packed, 1) if kind == "reference" else impl.decode(packed)
        t2 = time.perf_counter()
        assert dec is not None
        t_enc += t1 - t0
        t_dec += t2 - t1
    return t_enc, t_dec, len(seeds)


def cpu_reference_throughput(images_per_core: int):
    """One reference Encoder/Decoder(1 thread) per host core over disjoint images (the reference
    encoder is single threaded, encoder.cpp:59-109).  Returns (MP/s, cores, kind, sample)."""
    import multiprocessing as mp

    import oracle

    kind = "reference" if oracle.ref_available() else "port"
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    jobs = [(kind, [1 + c * images_per_core + i for i in range(images_per_core)]) for c in range(cores)]
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    n = sum(r[2] for r in res)
    mps = n * W * H / 1e6 / wall
    enc_core = n * W * H / 1e6 / sum(r[0] for r in res)
    dec_core = n * W * H / 1e6 / sum(r[1] for r in res)
    sample = (f"{n} 1080p RGB q50 images encode+decode, {images_per_core} per core on {cores} processes, "
              f"wall {wall:.1f}s; per-core encode {enc_core:.1f} MP/s, decode {dec_core:.1f} MP/s")
    return mps, cores, kind, sample


# ------------------------------------------------------------------------------------------------
def start_clock_sampler(path):
    q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        return subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    except OSError:
        return None


def summarize_clocks(path, gpu_index):
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    try:
        for line in open(path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9 or not f[0].isdigit() or int(f[0]) != gpu_index:
                continue
            sm.append(float(f[1]))
            mx.append(float(f[2]))
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
    except (OSError, ValueError):
        pass
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
    # median over the busier half of the samples (= under load)
    busy = sorted(sm)[len(sm) // 2:]
    return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--images", type=int, default=512, help="images per GPU (weak scaling)")
    ap.add_argument("--e2e-images", type=int, default=128, help="images per GPU for the host-buffer leg")
    ap.add_argument("--cpu-images-per-core", type=int, default=40)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    # ---- reference arm: the reference's CPU implementation, rank 0 only --------------------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        per = max(1, min(args.cpu_images_per_core, 12))  # bounded sample per step
        for _ in range(args.warmup):
            cpu_reference_throughput(1)
        vals, t0 = [], time.perf_counter()
        for _ in range(args.steps):
            vals.append(cpu_reference_throughput(per))
        wall = time.perf_counter() - t0
        mps = sum(v[0] for v in vals) / len(vals)
        _, cores, kind, sample = vals[-1]
        print(json.dumps({
            "impl": "reference", "metric": "encode+decode megapixels/sec", "value": mps, "unit": "MP/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "quality": QUALITY, "sample_images_per_step": per * cores},
            "cpu_baseline": {"value": mps, "unit": "MP/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": mps, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }))
        return 0

    # ---- CPU baseline first (fork before CUDA is initialised), rank 0 at N=1 only ---------------
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mps, cores, kind, sample = cpu_reference_throughput(max(1, args.cpu_images_per_core))
        cpu_baseline = {"value": mps, "unit": "MP/s", "cores": cores, "kind": kind, "sample": sample}

    import numpy as np
    import torch
    import torch.distributed as dist

    import himg_b200
    from himg_b200 import sharding
    from himg_b200.synth import synth_images

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = himg_b200.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    B = args.images
    total_images = B * world
    first, count = sharding.shard_range(total_images, world, rank)  # contiguous image ranges per GPU
    assert count == B
    pixels = synth_images(B, W, H, NCH, seed0=1 + first, amp=6, device=dev)
    out, sizes = ctx.encode_batch(pixels, QUALITY, True)
    offsets = torch.arange(B, dtype=torch.int64, device=dev) * out.stride(0)
    decoded, status = ctx.decode_batch(out.reshape(-1), offsets, sizes, W, H, NCH)
    torch.cuda.synchronize()
    assert int(status.abs().sum()) == 0 and int((sizes == 0).sum()) == 0

    def step():
        ctx.encode_batch(pixels, QUALITY, True, out=out, sizes=sizes)
        # the only collective of the job: per-image bitstream sizes -> global offset table
        table = sharding.gather_sizes(sizes, world)
        ctx.decode_batch(out.reshape(-1), offsets, sizes, W, H, NCH, out=decoded, status=status)
        return table

    for _ in range(warmup):
        step()
    barrier()

    clk_path = os.path.join(tempfile.gettempdir(), f"himg_clocks_{os.getpid()}.csv")
    sampler = start_clock_sampler(clk_path) if rank == 0 else None
    ctx.profile(True)
    ctx.profile_reset()
    launches0 = ctx.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    barrier()
    ev[0].record(stream)
    for _ in range(args.steps):
        table = step()
    ev[1].record(stream)
    barrier()
    ms = ev[0].elapsed_time(ev[1])
    launches = ctx.launch_count() - launches0
    prof = ctx.profile_results()
    ctx.profile(False)

    # encode-only / decode-only splits (device resident), same timing discipline
    def timed(fn, reps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        barrier()
        return a.elapsed_time(b) / reps

    enc_ms = timed(lambda: ctx.encode_batch(pixels, QUALITY, True, out=out, sizes=sizes), max(2, args.steps // 2))
    dec_ms = timed(lambda: ctx.decode_batch(out.reshape(-1), offsets, sizes, W, H, NCH, out=decoded, status=status),
                   max(2, args.steps // 2))

    # ---- e2e: host buffers through the C ABI, H2D/D2H inside the timed region ---------------------
    Be = min(args.e2e_images, B)
    h_pixels = torch.empty((Be, H, W, NCH), dtype=torch.uint8, pin_memory=True)
    h_pixels.copy_(pixels[:Be])
    bound = himg_b200.encode_bound(W, H, NCH)
    h_out = torch.empty((Be * bound,), dtype=torch.uint8, pin_memory=True)
    h_dec = torch.empty((Be, H, W, NCH), dtype=torch.uint8, pin_memory=True)
    h_off = np.zeros(Be + 1, np.uint64)
    h_sizes = np.zeros(Be, np.uint32)
    h_status = np.zeros(Be, np.int32)

    def e2e_step():
        ctx.encode_batch_host(h_pixels, QUALITY, True, out=h_out, offsets=h_off, sizes=h_sizes)
        ctx.decode_batch_host(h_out, h_off, h_sizes, W, H, NCH, out=h_dec, status=h_status)

    def wall(fn, reps):
        barrier()
        t0 = time.perf_counter()
        fn(reps)
        barrier()
        return (time.perf_counter() - t0) * 1e3 / reps

    for _ in range(2):
        e2e_step()
    assert int(np.abs(h_status).sum()) == 0
    assert torch.equal(h_dec[Be - 1], decoded[Be - 1].cpu()), "e2e leg decoded different pixels"
    e2e_steps = max(2, min(args.steps, 5))
    e2e_serial_ms = wall(lambda reps: [e2e_step() for _ in range(reps)], e2e_steps)

    # Pipelined steps: encode is H2D-heavy and decode D2H-heavy, so one host thread encodes step k+1 (its
    # own context) while another decodes step k: both PCIe directions stay busy.  Same public calls,
    # same bytes per step inside the timed region; two coding lanes per context measured best here.
    import queue
    import threading

    dev_index = dev.index if dev.index is not None else 0
    ctx_e, ctx_d = himg_b200.Context(dev_index), himg_b200.Context(dev_index)
    for c in (ctx_e, ctx_d):
        c.set_option("host_lanes", 2)
    h_out2 = [h_out, torch.empty(h_out.shape, dtype=torch.uint8, pin_memory=True)]
    h_off2 = [h_off, np.zeros(Be + 1, np.uint64)]
    h_sizes2 = [h_sizes, np.zeros(Be, np.uint32)]

    def e2e_pipelined(reps):
        ready, free = queue.SimpleQueue(), [threading.Semaphore(1), threading.Semaphore(1)]
        stop = threading.Event()

        def encode_leg():
            try:
                for k in range(reps):
                    b = k & 1
                    free[b].acquire()
                    if stop.is_set():
                        return
                    ctx_e.encode_batch_host(h_pixels, QUALITY, True, out=h_out2[b], offsets=h_off2[b], sizes=h_sizes2[b])
                    ready.put(b)
            except Exception as e:  # surface in the main thread
                ready.put(e)

        th = threading.Thread(target=encode_leg, daemon=True)
        th.start()
        try:
            for _ in range(reps):
                b = ready.get(timeout=120)
                if isinstance(b, Exception):
                    raise b
                ctx_d.decode_batch_host(h_out2[b], h_off2[b], h_sizes2[b], W, H, NCH, out=h_dec, status=h_status)
                free[b].release()
        finally:  # never leave the encode leg parked on a semaphore
            stop.set()
            for sem in free:
                sem.release()
            th.join(timeout=120)

    h_dec.zero_()
    e2e_pipelined(2)
    assert int(np.abs(h_status).sum()) == 0
    assert torch.equal(h_dec[Be - 1], decoded[Be - 1].cpu()), "pipelined e2e leg decoded different pixels"
    e2e_pipe_steps = max(e2e_steps, 8)  # the first encode and the last decode run alone: amortise them
    e2e_ms = wall(e2e_pipelined, e2e_pipe_steps)
    ctx_e.close()
    ctx_d.close()
    packed_bytes = int(h_off[Be])
    h2d = Be * W * H * NCH + packed_bytes
    d2h = packed_bytes + Be * W * H * NCH

    if sampler is not None:
        sampler.terminate()
        sampler.wait()

    # ---- max over ranks -------------------------------------------------------------------------
    def rmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def rsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms, enc_ms, dec_ms, e2e_ms, e2e_serial_ms = rmax(ms), rmax(enc_ms), rmax(dec_ms), rmax(e2e_ms), rmax(e2e_serial_ms)
    launches = int(rsum(launches))
    mp_per_step = total_images * W * H / 1e6
    value = mp_per_step * args.steps / (ms / 1e3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md 6.65 TB/s)"
        fwd_ms, fwd_n = prof.get("k_forward", (0.0, 0))
        inv_ms, inv_n = prof.get("k_inverse", (0.0, 0))
        alg_bytes = 2 * NCH * W * H * B  # per launch: nch read + nch written per pixel
        fwd_gbs = alg_bytes / (fwd_ms / fwd_n / 1e3) / 1e9 if fwd_n else 0.0
        inv_gbs = alg_bytes / (inv_ms / inv_n / 1e3) / 1e9 if inv_n else 0.0
        traffic = None  # DRAM bytes per launch from the committed ncu capture of this exact configuration
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_final_traffic.json")))
            if tr["config"] == {"images": B, "width": W, "height": H, "channels": NCH, "quality": QUALITY}:
                traffic = tr["k_forward"]["traffic"]
        except (OSError, ValueError, KeyError):
            pass
        step_kernel_ms = sum(v[0] for v in prof.values()) / args.steps
        shares = {k: round(v[0] / args.steps / step_kernel_ms, 4) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}
        res = {
            "metric": "encode+decode megapixels/sec",
            "value": value,
            "unit": "MP/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": warmup,
            "ms_per_step": ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "int16",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "images_per_gpu": B, "total_images": total_images, "quality": QUALITY,
                       "l2_policy": "inputs larger than L2 (3.2 GB of pixels per GPU per step)",
                       "parallelism": f"image-sharded x{world}, NCCL all_gather of bitstream sizes only"},
            "encode_mps": mp_per_step / (enc_ms / 1e3),
            "decode_mps": mp_per_step / (dec_ms / 1e3),
            "e2e": {"value": Be * world * W * H / 1e6 / (e2e_ms / 1e3), "unit": "MP/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "images_per_gpu": Be,
                    "steps": e2e_pipe_steps, "serial_value": Be * world * W * H / 1e6 / (e2e_serial_ms / 1e3),
                    "serial_steps": e2e_steps,
                    "api": "himgcu_encode_batch_host + himgcu_decode_batch_host on pinned host buffers (each call pipelines "
                           "H2D / coding lanes / D2H over sub-batches); value = consecutive steps pipelined by two host "
                           "threads (encode of step k+1 overlaps decode of step k: both PCIe directions busy); "
                           "serial_value = the two calls back to back, one step at a time"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "k_forward", "achieved": fwd_gbs, "peak": peak, "unit": "GB/s",
                         "frac": fwd_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": fwd_ms / fwd_n if fwd_n else None,
                         "k_inverse": {"achieved": inv_gbs, "frac": inv_gbs / peak,
                                       "avg_launch_ms": inv_ms / inv_n if inv_n else None}},
            "kernel_time_shares": shares,
            "kernel_ms_per_step": step_kernel_ms,
            "cpu_baseline": cpu_baseline,
            "clocks": summarize_clocks(clk_path, local_rank),
            "packed_bits_per_pixel": 8.0 * packed_bytes / (Be * W * H),
        }
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
