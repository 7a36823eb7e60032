"""BASELINE config c5: quality sweep on a 1080p RGB batch (device resident).  Prints a table of
encode / decode MP/s, bits per pixel and strict-decode acceptance per quality."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import himg_b200
from himg_b200.synth import synth_images

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
W, H, N = 1920, 1080, 3
ctx = himg_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
px = synth_images(B, W, H, N, 1, 6)
rows = []
for q in range(0, 101, 10):
    out, sizes = ctx.encode_batch(px, q, True)
    offs = torch.arange(B, dtype=torch.int64, device="cuda") * out.stride(0)
    dec, st = ctx.decode_batch(out.reshape(-1), offs, sizes, W, H, N, flags=0)
    strict_ok = int((st == 0).sum())
    dec, st = ctx.decode_batch(out.reshape(-1), offs, sizes, W, H, N, flags=1)
    assert int(st.abs().sum()) == 0
    def timed(fn, reps=3):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    e = timed(lambda: ctx.encode_batch(px, q, True, out=out, sizes=sizes))
    d = timed(lambda: ctx.decode_batch(out.reshape(-1), offs, sizes, W, H, N, flags=1, out=dec, status=st))
    mp = B * W * H / 1e6
    rows.append({"quality": q, "encode_mps": round(mp / e * 1e3), "decode_mps": round(mp / d * 1e3),
                 "bpp": round(float(sizes.sum()) * 8 / (B * W * H), 3), "strict_decode_accepts": f"{strict_ok}/{B}"})
    print(rows[-1])
json.dump(rows, open("gpurun_out/quality_sweep.json", "w"), indent=1)
