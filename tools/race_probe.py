"""Batch case of gpu_smoke with details (which stage differs), for runs under compute-sanitizer."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import himg_b200  # noqa: E402
import oracle  # noqa: E402

P = oracle.port()
ctx = himg_b200.Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)  # torch prepares the inputs on this stream
B, w, h, n, q = int(sys.argv[1]) if len(sys.argv) > 1 else 136, 256, 256, 3, 50
imgs = np.stack([P.synth(w, h, n, 100 + k, 6) for k in range(B)])
out, sizes = ctx.encode_batch(torch.from_numpy(imgs).cuda(), q, True)
offs = torch.arange(B, dtype=torch.int64, device="cuda") * out.stride(0)
dec, status = ctx.decode_batch(out.reshape(-1), offs, sizes, w, h, n)
out, sizes, dec, status = out.cpu().numpy(), sizes.cpu().numpy(), dec.cpu().numpy(), status.cpu().numpy()
print("status nonzero:", np.flatnonzero(status).tolist()[:10])
bad_e, bad_d = [], []
for k in range(B):
    want = P.encode(imgs[k], q, True)
    if bytes(out[k, : sizes[k]]) != want:
        a = np.frombuffer(want, np.uint8)
        b = out[k, : sizes[k]]
        m = min(a.size, b.size)
        d = np.flatnonzero(a[:m] != b[:m])
        bad_e.append((k, int(sizes[k]), len(want), int(d[0]) if d.size else -1, int(d.size)))
    wd = P.decode(want, strict=False)
    if not np.array_equal(dec[k], wd):
        dd = np.argwhere(dec[k] != wd)
        bad_d.append((k, len(dd), dd[0].tolist(), dd[-1].tolist()))
print("encode mismatches:", bad_e[:8])
print("decode mismatches:", bad_d[:8])
