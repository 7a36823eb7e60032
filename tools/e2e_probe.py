import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import himg_b200
from himg_b200.synth import synth_images
W, H, N, B = 1920, 1080, 3, 128
ctx = himg_b200.Context(0)
px = synth_images(B, W, H, N, 1, 6)
h_px = torch.empty((B, H, W, N), dtype=torch.uint8, pin_memory=True); h_px.copy_(px)
d = torch.empty_like(px)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
ms = t(lambda: d.copy_(h_px, non_blocking=True)); print(f"H2D {h_px.numel()/1e6:.0f} MB: {ms:.2f} ms = {h_px.numel()/ms/1e6:.1f} GB/s")
h2 = torch.empty_like(h_px).pin_memory()
ms = t(lambda: h2.copy_(d, non_blocking=True)); print(f"D2H: {ms:.2f} ms = {h_px.numel()/ms/1e6:.1f} GB/s")
bound = himg_b200.encode_bound(W, H, N)
h_out = torch.empty((B * bound,), dtype=torch.uint8, pin_memory=True)
off = np.zeros(B + 1, np.uint64); sz = np.zeros(B, np.uint32); st = np.zeros(B, np.int32)
ms = t(lambda: ctx.encode_batch_host(h_px, 50, True, out=h_out, offsets=off, sizes=sz)); print(f"encode_batch_host: {ms:.2f} ms")
ms = t(lambda: ctx.decode_batch_host(h_out, off, sz, W, H, N, out=h2, status=st)); print(f"decode_batch_host: {ms:.2f} ms")
for sub in (32, 64, 128, 192, 256, 384, 512):
    ctx.set_option("host_sub_batch_bytes", sub << 20)
    ms1 = t(lambda: ctx.encode_batch_host(h_px, 50, True, out=h_out, offsets=off, sizes=sz))
    ms2 = t(lambda: ctx.decode_batch_host(h_out, off, sz, W, H, N, out=h2, status=st))
    print(f"sub {sub} MB: encode {ms1:.2f} ms decode {ms2:.2f} ms")
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
out, sizes = ctx.encode_batch(px, 50, True)
ms = t(lambda: ctx.encode_batch(px, 50, True, out=out, sizes=sizes)); print(f"device encode: {ms:.2f} ms")
