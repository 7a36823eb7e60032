"""Encode leg and decode leg of the host-buffer calls: alone, then concurrently from two threads
(two contexts, both PCIe directions busy).  usage: e2e_duplex_probe.py [chunks] [sub_batch_MB] [lanes];
TRACE_ONLY=1 prints the device-side timeline of one concurrent run (HIMG_DEBUG_PIPE)."""
import sys, time, threading
sys.path.insert(0, ".")
import numpy as np, torch
import himg_b200
from himg_b200.synth import synth_images

W, H, NCH, Q, B, CH = 1920, 1080, 3, 50, int(__import__("os").environ.get("PROBE_B", "128")), int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
import os
ctx_e, ctx_d = himg_b200.Context(0), himg_b200.Context(0)
SUB = int(sys.argv[2]) << 20 if len(sys.argv) > 2 else 64 << 20
LANES = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx_e.set_option("host_lanes", LANES); ctx_d.set_option("host_lanes", LANES)
ctx_e.set_option("host_sub_batch_bytes", SUB); ctx_d.set_option("host_sub_batch_bytes", SUB)
px = synth_images(B, W, H, NCH, seed0=1, amp=6, device=dev)
h_px = torch.empty((B, H, W, NCH), dtype=torch.uint8, pin_memory=True); h_px.copy_(px)
bound = himg_b200.encode_bound(W, H, NCH)
h_out = torch.empty((B * bound,), dtype=torch.uint8, pin_memory=True)
h_dec = torch.empty((B, H, W, NCH), dtype=torch.uint8, pin_memory=True)
n = B // CH
offs = [np.zeros(n + 1, np.uint64) for _ in range(CH)]
sizes = np.zeros(B, np.uint32); status = np.zeros(B, np.int32)

def enc():
    for k in range(CH):
        a, b = k * n, (k + 1) * n
        ctx_e.encode_batch_host(h_px[a:b], Q, True, out=h_out[a * bound:b * bound], offsets=offs[k], sizes=sizes[a:b])

def dec():
    for k in range(CH):
        a, b = k * n, (k + 1) * n
        ctx_d.decode_batch_host(h_out[a * bound:b * bound], offs[k], sizes[a:b], W, H, NCH, out=h_dec[a:b], status=status[a:b])

def both():
    t = threading.Thread(target=enc); t.start(); dec(); t.join()

def timeit(fn, reps=4):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps

enc(); dec()
import os
if os.environ.get("TRACE_ONLY"):
    os.environ["HIMG_DEBUG_PIPE"] = "1"; torch.cuda.synchronize(); both(); sys.exit(0)
print(f"lanes={LANES} sub={SUB>>20}MB chunks={CH}: enc {timeit(enc):.1f} ms  dec {timeit(dec):.1f} ms  both(threads) {timeit(both):.1f} ms")
