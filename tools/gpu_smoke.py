"""Tiny encode+decode through the C ABI, checked against the oracle (used under compute-sanitizer)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import himg_b200  # noqa: E402
import oracle  # noqa: E402


def main():
    P = oracle.port()
    ctx = himg_b200.Context(0)
    ok = True
    # the last two shapes have a low-res chunk large enough for the multi-CTA packer (parts) and the
    # cluster decoder
    for (w, h, n, q) in [(64, 48, 3, 50), (37, 21, 1, 60), (200, 136, 4, 90), (264, 40, 3, 20), (512, 64, 3, 50), (8192, 16, 1, 100),
                         (256, 128, 4, 80), (1024, 1024, 3, 50), (2048, 1024, 3, 70), (72, 40, 6, 50)]:
        img = P.synth(w, h, n, 3, 6)
        got = ctx.encode(img, q, True)
        want = P.encode(img, q, True)
        e_ok = got == want
        dec = ctx.decode(want, 1)
        wd = P.decode(want, strict=False)
        d_ok = dec is not None and np.array_equal(dec, wd)
        print(f"{w}x{h}x{n} q{q}: encode {'OK' if e_ok else 'MISMATCH'} ({len(got)} vs {len(want)}), decode {'OK' if d_ok else 'MISMATCH'}")
        ok &= e_ok and d_ok
    # a batch with enough block rows for the warp-team stream decoder (>= 4096 streams) and the
    # lane-pair inverse transform
    import torch

    ctx.set_stream(torch.cuda.current_stream().cuda_stream)  # torch prepares the inputs on this stream
    B, w, h, n, q = 136, 256, 256, 3, 50
    imgs = np.stack([P.synth(w, h, n, 100 + k, 6) for k in range(B)])
    out, sizes = ctx.encode_batch(torch.from_numpy(imgs).cuda(), q, True)
    offs = torch.arange(B, dtype=torch.int64, device="cuda") * out.stride(0)
    dec, status = ctx.decode_batch(out.reshape(-1), offs, sizes, w, h, n)
    out, sizes, dec = out.cpu().numpy(), sizes.cpu().numpy(), dec.cpu().numpy()
    b_ok = int(status.abs().sum()) == 0
    for k in range(0, B, 17):
        want = P.encode(imgs[k], q, True)
        b_ok &= bytes(out[k, : sizes[k]]) == want and np.array_equal(dec[k], P.decode(want, strict=False))
    print(f"batch {B} x {w}x{h}x{n} q{q}: {'OK' if b_ok else 'MISMATCH'}")
    ok &= b_ok
    print("SMOKE", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
