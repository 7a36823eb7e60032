"""One single-image encode+decode (device resident) for ncu: python tools/prof_single.py W H NCH"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import himg_b200  # noqa: E402
from himg_b200.synth import synth_images  # noqa: E402

W, H, N = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (3840, 2160, 3)
ctx = himg_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
px = synth_images(1, W, H, N, 1, 6)
for _ in range(2):
    out, sizes = ctx.encode_batch(px, 50, True)
    offs = torch.zeros(1, dtype=torch.int64, device="cuda")
    dec, st = ctx.decode_batch(out.reshape(-1), offs, sizes, W, H, N)
torch.cuda.synchronize()
print("size", sizes.tolist(), "status", int(st.abs().sum()))
