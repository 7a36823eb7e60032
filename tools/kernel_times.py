"""Per-kernel event times (library profiling hooks) of one device-resident encode+decode batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from himg_b200 import _native  # noqa: E402

if os.environ.get("HIMG_AB_LIB"):  # A/B runs of this tool only: time another build of the library
    _native.LIB_PATH = os.environ["HIMG_AB_LIB"]
import himg_b200  # noqa: E402
from himg_b200.synth import synth_images  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 50
W, H, N = 1920, 1080, 3
ctx = himg_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
px = synth_images(B, W, H, N, 1, 6)


def step():
    out, sizes = ctx.encode_batch(px, Q, True)
    offs = torch.arange(B, dtype=torch.int64, device="cuda") * out.stride(0)
    return ctx.decode_batch(out.reshape(-1), offs, sizes, W, H, N)


for _ in range(3):
    step()
ctx.synchronize()
ctx.profile(True)
ctx.profile_reset()
reps = 4
for _ in range(reps):
    step()
ctx.synchronize()
res = ctx.profile_results()
tot = sum(v[0] for v in res.values())
print(f"# {B} images {W}x{H}x{N} q{Q}: {tot / reps:.3f} ms per step in kernels, {B * W * H / 1e6 / (tot / reps / 1e3):.0f} MP/s")
for k, (ms, cnt) in sorted(res.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:22s} {ms / reps:8.3f} ms  {ms / tot * 100:5.1f}%  launches/step={cnt // reps}")
