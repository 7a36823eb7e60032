"""Generate the golden vectors under tests/golden/ from the UNMODIFIED reference.

Run in the build container (needs /root/reference, compiled in place into oracle/_ref by
oracle/build_oracle.py):   python tests/golden/make_golden.py

    python tests/golden/make_golden.py --shards     (only golden_c4_shards.json)

Outputs (committed):
  golden_c4_shards.json  the same hashes for the first and the last image of every rank's shard of the
                       c4 batch (4096 1080p images, seeds 1..4096, quality 50) on 1, 2, 4 and 8 GPUs: what
                       bench.py checks inside every run, on every rank.
  golden_hashes.json   FNV-1a-64 hashes + sizes of reference .himg output / decoded pixels for the
                       BASELINE.json configurations and edge cases (SURVEY Appendix B generator).
  fixtures.npz         a few small complete reference bitstreams + stage-level vectors.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (w, h, nch, quality, seed, amp, use_ycbcr)
HASH_CASES = [
    (512, 512, 3, 50, 1, 6, 1),      # c1
    (3840, 2160, 3, 50, 1, 6, 1),    # c2
    (8192, 8192, 1, 50, 1, 6, 1),    # c3
] + [(1920, 1080, 3, q, 1, 6, 1) for q in range(0, 101, 10)] + [  # c4 / c5
    (1920, 1080, 3, 50, 2, 6, 1),
    (1920, 1080, 3, 50, 4096, 6, 1),
    (512, 512, 3, 20, 7, 6, 1),
    (512, 512, 1, 90, 7, 6, 1),
    (8192, 16, 1, 100, 7, 6, 1),     # 4-byte segment headers
    (256, 256, 4, 75, 7, 6, 1),
    (512, 512, 3, 100, 7, 6, 1),
    (512, 512, 3, 50, 7, 6, 0),      # -rgb
    (500, 300, 3, 50, 3, 6, 1),      # width % 8 != 0 (reference decode is UB: encode only)
    (37, 21, 1, 60, 3, 6, 1),
    (512, 300, 3, 50, 3, 6, 1),      # height % 8 != 0
    (64, 8, 3, 50, 3, 6, 1),         # single block row
    (8, 8, 1, 50, 3, 6, 1),
    (1, 1, 3, 50, 3, 6, 1),
    (640, 480, 3, 50, 5, 0, 1),      # noise-free
    (640, 480, 3, 95, 5, 40, 1),     # noisy
]

# small complete bitstreams kept verbatim
FIXTURE_CASES = [
    (64, 48, 3, 50, 11, 6, 1),
    (40, 24, 1, 60, 12, 6, 1),
    (37, 21, 1, 60, 3, 6, 1),
    (96, 64, 4, 75, 13, 6, 1),
    (128, 64, 3, 100, 14, 20, 1),
    (128, 64, 3, 5, 15, 6, 1),
    (72, 40, 3, 50, 16, 6, 0),
]


def shard_seeds(total=4096):
    seeds = set()
    for world in (1, 2, 4, 8):
        per = total // world
        for r in range(world):
            seeds.update((1 + r * per, (r + 1) * per))
    return sorted(seeds)


def shard_hashes():
    P, R = oracle.port(), oracle.ref()
    out = []
    for seed in shard_seeds():
        img = P.synth(1920, 1080, 3, seed, 6)
        packed = R.encode(img, 50, True)
        dec = R.decode(packed, 1)
        assert dec is not None
        out.append({"w": 1920, "h": 1080, "nch": 3, "quality": 50, "seed": seed, "amp": 6, "ycbcr": 1,
                    "himg_size": len(packed), "himg_hash": f"{P.fnv(np.frombuffer(packed, np.uint8)):016x}",
                    "pixel_hash": f"{P.fnv(dec):016x}"})
        print(out[-1])
    with open(os.path.join(HERE, "golden_c4_shards.json"), "w") as f:
        json.dump(out, f, indent=1)


def main():
    if "--shards" in sys.argv:
        return shard_hashes()
    P, R = oracle.port(), oracle.ref()
    hashes = []
    for (w, h, n, q, seed, amp, yc) in HASH_CASES:
        img = P.synth(w, h, n, seed, amp)
        packed = R.encode(img, q, bool(yc))
        dec = R.decode(packed) if w % 8 == 0 else None  # UB in the reference otherwise
        hashes.append({
            "w": w, "h": h, "nch": n, "quality": q, "seed": seed, "amp": amp, "ycbcr": yc,
            "src_hash": f"{P.fnv(img):016x}",
            "himg_size": len(packed),
            "himg_hash": f"{P.fnv(np.frombuffer(packed, np.uint8)):016x}",
            "ref_decode": "ok" if dec is not None else ("skipped" if w % 8 else "fails"),
            "pixel_hash": f"{P.fnv(dec):016x}" if dec is not None else None,
        })
        print(hashes[-1])
    with open(os.path.join(HERE, "golden_hashes.json"), "w") as f:
        json.dump(hashes, f, indent=1)

    fx = {}
    for i, (w, h, n, q, seed, amp, yc) in enumerate(FIXTURE_CASES):
        img = P.synth(w, h, n, seed, amp)
        packed = R.encode(img, q, bool(yc))
        fx[f"case{i}_meta"] = np.array([w, h, n, q, seed, amp, yc], np.int32)
        fx[f"case{i}_himg"] = np.frombuffer(packed, np.uint8)
        if w % 8 == 0:
            dec = R.decode(packed)
            if dec is not None:
                fx[f"case{i}_pixels"] = dec
    # stage-level: Huffman chunk on structured random data (framed + unframed)
    rng = np.random.default_rng(1234)
    data = rng.integers(0, 256, 4096, dtype=np.uint8)
    data[rng.random(4096) < 0.7] = 0
    data[1000:1400] = 0
    fx["huff_in"] = data
    fx["huff_unframed"] = np.frombuffer(R.huff_compress(data, 0), np.uint8)
    fx["huff_framed512"] = np.frombuffer(R.huff_compress(data, 512), np.uint8)
    # stage-level: low-res of one channel
    img = P.synth(200, 136, 3, 21, 6)
    cm = R.rgb_to_ycbcr(img)
    L, bd = R.lowres_channel(cm, 0, 50)
    fx["lowres_src"] = img
    fx["lowres_L_ch0"] = L
    fx["lowres_blockdata_ch0"] = bd
    # stage-level: quantise/map + transforms on random blocks
    blk = rng.integers(-255, 256, (8, 64)).astype(np.int16)
    fx["wht_in"] = blk
    fx["wht_fwd"] = np.stack([R.hadamard_forward(b) for b in blk])
    coef = rng.integers(-16320, 16321, (8, 64)).astype(np.int16)
    fx["quant_in"] = coef
    fx["quant_q50_luma"] = np.stack([R.quant_pack(50, 0, b) for b in coef])
    fx["quant_q80_chroma"] = np.stack([R.quant_pack(80, 1, b) for b in coef])
    codes = rng.integers(0, 256, (8, 64)).astype(np.uint8)
    codes[codes == 128] = 0
    fx["dequant_in"] = codes
    fx["dequant_q50_luma"] = np.stack([R.quant_unpack(50, 0, b) for b in codes])
    fx["wht_inv"] = np.stack([R.hadamard_inverse(b) for b in fx["dequant_q50_luma"]])
    fx["qcfg_q50"] = np.frombuffer(R.quant_config(50, 1), np.uint8)
    fx["qcfg_q0"] = np.frombuffer(R.quant_config(0, 1), np.uint8)
    fx["qcfg_q100"] = np.frombuffer(R.quant_config(100, 1), np.uint8)
    fx["qcfg_q73"] = np.frombuffer(R.quant_config(73, 1), np.uint8)
    fx["lmap_q50"] = np.frombuffer(R.lowres_mapfun(50), np.uint8)
    fx["lmap_q0"] = np.frombuffer(R.lowres_mapfun(0), np.uint8)
    fx["lmap_q7"] = np.frombuffer(R.lowres_mapfun(7), np.uint8)
    fx["fmap"] = np.frombuffer(R.fullres_mapfun(), np.uint8)
    np.savez_compressed(os.path.join(HERE, "fixtures.npz"), **fx)
    print("fixtures:", os.path.getsize(os.path.join(HERE, "fixtures.npz")), "bytes")


if __name__ == "__main__":
    main()
