"""Short device-resident encode+decode run for profiling under ncu (never a bench number)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import himg_b200  # noqa: E402
from himg_b200.synth import synth_images  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
W, H, N = 1920, 1080, 3
ctx = himg_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
px = synth_images(B, W, H, N, 1, 6)
for _ in range(reps):
    out, sizes = ctx.encode_batch(px, 50, True)
    offs = torch.arange(B, dtype=torch.int64, device="cuda") * out.stride(0)
    dec, st = ctx.decode_batch(out.reshape(-1), offs, sizes, W, H, N)
torch.cuda.synchronize()
print("sizes", sizes[:3].tolist(), "status", int(st.abs().sum()))
