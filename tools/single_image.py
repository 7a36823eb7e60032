"""Latency of the single-image BASELINE configs (c1..c3) through the device batch API (n=1) and the
host API, with the per-kernel breakdown.  Not a bench line; numbers go to DESIGN.md."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import himg_b200
from himg_b200.synth import synth_images

ctx = himg_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
res = {}
for name, (w, h, n) in {"c1_512x512x3": (512, 512, 3), "c2_3840x2160x3": (3840, 2160, 3), "c3_8192x8192x1": (8192, 8192, 1)}.items():
    px = synth_images(1, w, h, n, 1, 6)
    out, sizes = ctx.encode_batch(px, 50, True)
    offs = torch.zeros(1, dtype=torch.int64, device="cuda")
    dec, st = ctx.decode_batch(out.reshape(-1), offs, sizes, w, h, n)
    torch.cuda.synchronize()
    assert int(st[0]) == 0
    def timed(fn, reps=20):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    enc_ms = timed(lambda: ctx.encode_batch(px, 50, True, out=out, sizes=sizes))
    dec_ms = timed(lambda: ctx.decode_batch(out.reshape(-1), offs, sizes, w, h, n, out=dec, status=st))
    ctx.profile(True); ctx.profile_reset()
    for _ in range(5):
        ctx.encode_batch(px, 50, True, out=out, sizes=sizes)
        ctx.decode_batch(out.reshape(-1), offs, sizes, w, h, n, out=dec, status=st)
    prof = {k: round(v[0] / 5 * 1e3, 1) for k, v in sorted(ctx.profile_results().items(), key=lambda kv: -kv[1][0])}
    ctx.profile(False)
    img = px[0].cpu().numpy()
    t0 = time.perf_counter(); packed = ctx.encode(img, 50, True); t1 = time.perf_counter(); back = ctx.decode(packed); t2 = time.perf_counter()
    t0 = time.perf_counter(); packed = ctx.encode(img, 50, True); t1 = time.perf_counter(); back = ctx.decode(packed); t2 = time.perf_counter()
    mp = w * h / 1e6
    res[name] = {"device_encode_ms": round(enc_ms, 3), "device_decode_ms": round(dec_ms, 3), "encode_mps": round(mp / enc_ms * 1e3), "decode_mps": round(mp / dec_ms * 1e3),
                 "host_api_encode_ms": round((t1 - t0) * 1e3, 2), "host_api_decode_ms": round((t2 - t1) * 1e3, 2), "bytes": len(packed), "kernel_us": prof}
    print(name, json.dumps(res[name]))
json.dump(res, open("gpurun_out/single_image.json", "w"), indent=1)
