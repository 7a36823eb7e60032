"""K-fwd / K-inv alone (stage entry points), CUDA-event times and % of the HBM roofline, per kernel variant.

    python tools/xform_bench.py [images] [variants, e.g. 0,1]

Algorithmic bytes = 2 * nch per pixel (read nch + write nch); peak from MEASURED_PEAKS.json.  The outputs of
all variants are compared with each other (the parity tests compare them with the oracle)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import himg_b200  # noqa: E402
from himg_b200.synth import synth_images  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
VARIANTS = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,1").split(",")]
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except (OSError, ValueError, KeyError):
    PEAK = 6650.0

CASES = [("c4 1080p RGB", 1920, 1080, 3, B), ("c2 4K RGB", 3840, 2160, 3, 1), ("c3 8K gray", 8192, 8192, 1, 1),
         ("720p RGB", 1280, 720, 3, max(1, B // 2))]


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / reps)
    return best


def main():
    ctx = himg_b200.Context(0)
    ctx.set_stream(torch.cuda.current_stream().cuda_stream)
    res = {}
    for name, w, h, nch, n in CASES:
        px = synth_images(n, w, h, nch, 1, 6)
        yc = nch >= 3
        L = ctx.stage_lowres(px, yc)
        out, sizes = ctx.encode_batch(px, 50, yc)
        # decoder-side inputs: planes from the encoder's own forward stage, tables of quality 50
        import oracle

        port = oracle.port()
        unmap = port.mapfun_parse(port.mapfun_serialize(port.fullres_map_table()))
        sl, sc = port.shift_table(50, 0), port.shift_table(50, 1)
        ref_planes = ref_pixels = None
        alg = 2 * nch * w * h * n
        for v in VARIANTS:
            ctx.set_option("xform_variant", v)
            planes = ctx.stage_forward(px, L, 50, yc)
            pixels = ctx.stage_inverse(planes, L, w, h, nch, yc, sl, sc, unmap)
            if ref_planes is None:
                ref_planes, ref_pixels = planes, pixels
            else:
                assert torch.equal(planes, ref_planes), f"{name}: forward variants differ"
                assert torch.equal(pixels, ref_pixels), f"{name}: inverse variants differ"
            reps = 5 if n > 1 else 50
            t_f = timed(lambda: ctx.stage_forward(px, L, 50, yc), reps)
            t_i = timed(lambda: ctx.stage_inverse(planes, L, w, h, nch, yc, sl, sc, unmap), reps)
            res[f"{name} v{v}"] = {"images": n, "fwd_ms": t_f, "fwd_frac": alg / t_f / 1e6 / PEAK, "inv_ms": t_i,
                                   "inv_frac": alg / t_i / 1e6 / PEAK}
            print(f"{name:14s} variant {v}: K-fwd {t_f * 1e3:9.1f} us = {alg / t_f / 1e6:7.0f} GB/s = {alg / t_f / 1e6 / PEAK * 100:5.1f}%   "
                  f"K-inv {t_i * 1e3:9.1f} us = {alg / t_i / 1e6:7.0f} GB/s = {alg / t_i / 1e6 / PEAK * 100:5.1f}%", flush=True)
        del px, L, out, planes, pixels, ref_planes, ref_pixels
        torch.cuda.empty_cache()
    ctx.set_option("xform_variant", 0)
    print(json.dumps({"peak_gbs": PEAK, "results": res}))


if __name__ == "__main__":
    main()
