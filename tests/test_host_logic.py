"""CPU tests of the host-side logic: C-ABI symbols, sharding (gloo, world_size 2), generator."""
import ctypes
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """libhimgcu.so loads without a GPU and exports exactly what include/himg_cuda.h declares."""
    from himg_b200 import _native
    from himg_b200 import build as hb

    hb.build()
    lib = _native.load()
    header = open(os.path.join(ROOT, "include", "himg_cuda.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(himgcu_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared == set(_native.SIGNATURES), (declared ^ set(_native.SIGNATURES))
    for name in declared:
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr)
    assert lib.himgcu_abi_version() == 1


def test_no_cpu_fallback():
    """Without a device the product path fails loudly instead of computing on the CPU."""
    import torch

    import himg_b200

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(himg_b200.HimgError):
        himg_b200.Context(0)
    assert himg_b200.encode_bound(1920, 1080, 3) > 6220800


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "himg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "himg_oracle" not in src, f


def test_encode_bound_and_lres_sizes():
    from himg_b200 import _native

    lib = _native.load()
    for (w, h, n, lres) in [(512, 512, 3, 12336), (3840, 2160, 3, 390330), (8192, 8192, 1, 1052672), (1920, 1080, 3, 97605)]:
        assert lib.himgcu_lres_size(w, h, n) == lres
        assert lib.himgcu_lres_stride(w, h, n) % 64 == 0 and lib.himgcu_lres_stride(w, h, n) >= lres
        assert lib.himgcu_encode_bound(w, h, n) > w * h * n
    assert lib.himgcu_encode_bound(0, 10, 3) == 0 and lib.himgcu_encode_bound(10, 10, 256) == 0
    assert lib.himgcu_encode_bound(10, 10, 5) > 500  # any channel count up to 255, as the reference


def test_decode_info_host_only(port):
    from himg_b200 import _native

    lib = _native.load()
    packed = np.frombuffer(port.encode(port.synth(72, 40, 3, 1, 6), 50, True), np.uint8)
    w, h, n = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.himgcu_decode_info(packed.ctypes.data, packed.size, ctypes.byref(w), ctypes.byref(h), ctypes.byref(n)) == 0
    assert (w.value, h.value, n.value) == (72, 40, 3)
    assert lib.himgcu_decode_info(packed.ctypes.data, packed.size - 1, ctypes.byref(w), ctypes.byref(h), ctypes.byref(n)) == 1


def test_shard_range_partitions():
    from himg_b200.sharding import shard_range

    for total in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            got = [shard_range(total, world, r) for r in range(world)]
            assert got[0][0] == 0 and sum(c for _, c in got) == total
            for (f0, c0), (f1, _) in zip(got, got[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in got) - min(c for _, c in got) <= 1


def test_gather_sizes_single_rank():
    import torch

    from himg_b200.sharding import gather_sizes

    sizes = torch.tensor([100, 33, 16, 1], dtype=torch.int32)
    allsz, off = gather_sizes(sizes, 1)
    assert allsz.tolist() == [100, 33, 16, 1] and off.tolist() == [0, 112, 160, 176, 192]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from himg_b200.sharding import gather_sizes, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 11
    first, count = shard_range(total, world, rank)
    sizes = torch.tensor([1000 + 37 * (first + i) for i in range(count)], dtype=torch.int32)
    allsz, off = gather_sizes(sizes)
    q.put((rank, first, count, allsz.tolist(), off.tolist()))
    dist.destroy_process_group()


def test_gather_sizes_gloo_world2():
    """The N>1 path on CPU: two gloo ranks with unequal shards end up with the same global table,
    identical to the single-process table."""
    import torch.multiprocessing as mp

    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want_sizes = [1000 + 37 * i for i in range(11)]
    want_off = [0]
    for s in want_sizes:
        want_off.append(want_off[-1] + (s + 15) // 16 * 16)
    assert res[0][1:3] == (0, 6) and res[1][1:3] == (6, 5)
    for _, _, _, allsz, off in res:
        assert allsz == want_sizes and off == want_off


def test_synth_generator_matches_oracle(port):
    import torch  # noqa: F401

    from himg_b200.synth import synth_images

    for (w, h, n, s, a) in [(64, 48, 3, 1, 6), (200, 100, 1, 77, 0), (128, 72, 4, 4096, 40)]:
        g = synth_images(2, w, h, n, s, a, device="cpu").numpy()
        assert np.array_equal(g[0], port.synth(w, h, n, s, a))
        assert np.array_equal(g[1], port.synth(w, h, n, s + 1, a))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with
    the contract's keys; a one-image-per-core sample keeps this to a few seconds."""
    import json
    import subprocess

    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-images-per-core", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "encode+decode megapixels/sec" and d["unit"] == "MP/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
