"""Host-buffer calls on 512 1080p images: encode leg alone, decode leg alone, both at once from two threads
(what bench.py's e2e leg does), for several lane counts and sub-batch sizes."""
import sys, time, threading
sys.path.insert(0, ".")
import numpy as np, torch
import himg_b200
from himg_b200.synth import synth_images

W, H, NCH, Q, B = 1920, 1080, 3, 50, 512
dev = torch.device("cuda:0")
px = synth_images(B, W, H, NCH, seed0=1, amp=6, device=dev)
h_px = torch.empty((B, H, W, NCH), dtype=torch.uint8, pin_memory=True); h_px.copy_(px)
del px
bound = himg_b200.encode_bound(W, H, NCH)
h_out = torch.empty((B * bound,), dtype=torch.uint8, pin_memory=True)
h_dec = torch.empty((B, H, W, NCH), dtype=torch.uint8, pin_memory=True)
off = np.zeros(B + 1, np.uint64); sizes = np.zeros(B, np.uint32); status = np.zeros(B, np.int32)
MP = B * W * H / 1e6

def run(lanes, sub_mb):
    ctx_e, ctx_d = himg_b200.Context(0), himg_b200.Context(0)
    for c in (ctx_e, ctx_d):
        c.set_option("host_lanes", lanes); c.set_option("host_sub_batch_bytes", sub_mb << 20)
    enc = lambda: ctx_e.encode_batch_host(h_px, Q, True, out=h_out, offsets=off, sizes=sizes)
    dec = lambda: ctx_d.decode_batch_host(h_out, off, sizes, W, H, NCH, out=h_dec, status=status)
    def both():
        t = threading.Thread(target=enc); t.start(); dec(); t.join()
    def timeit(fn, reps=3):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / reps
    e, d, b = timeit(enc), timeit(dec), timeit(both)
    print(f"lanes={lanes} sub={sub_mb:4d} MB: enc {e:6.1f} ms  dec {d:6.1f} ms  both {b:6.1f} ms  -> pipelined {MP / b * 1e3:7.0f} MP/s, serial {MP / (e + d) * 1e3:7.0f} MP/s", flush=True)
    ctx_e.close(); ctx_d.close()

for lanes in (3, 4):
    for sub in (128, 192, 256):
        run(lanes, sub)
