"""K-fwd / K-inv timing on the bench workload (stage calls, CUDA events)."""
import sys
sys.path.insert(0, ".")
import torch
import himg_b200
from himg_b200.synth import synth_images

W, H, NCH, Q, B = 1920, 1080, 3, 50, int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda:0")
ctx = himg_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream())
px = synth_images(B, W, H, NCH, seed0=1, amp=6, device=dev)
L = ctx.stage_lowres(px, True)
planes = ctx.stage_forward(px, L, Q, True)
ref = planes.clone()


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ms = timed(lambda: ctx.stage_forward(px, L, Q, True))
gb = 2 * B * W * H * NCH / 1e9
print(f"k_forward: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  ({gb / ms * 1e3 / 6461.5 * 100:.1f}% of 6461.5)  checksum {int(ref.to(torch.int64).sum())}")
