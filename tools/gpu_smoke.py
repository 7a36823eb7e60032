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
    for (w, h, n, q) in [(64, 48, 3, 50), (37, 21, 1, 60), (200, 136, 4, 90), (264, 40, 3, 20), (512, 64, 3, 50), (8192, 16, 1, 100), (256, 128, 4, 80)]:
        img = P.synth(w, h, n, 3, 6)
        got = ctx.encode(img, q, True)
        want = P.encode(img, q, True)
        e_ok = got == want
        dec = ctx.decode(want, 1)
        wd = P.decode(want, strict=False)
        d_ok = dec is not None and np.array_equal(dec, wd)
        print(f"{w}x{h}x{n} q{q}: encode {'OK' if e_ok else 'MISMATCH'} ({len(got)} vs {len(want)}), decode {'OK' if d_ok else 'MISMATCH'}")
        ok &= e_ok and d_ok
    print("SMOKE", "PASS" if ok else "FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
