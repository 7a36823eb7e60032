"""The C++ host layer (himg::Encoder / himg::Decoder in libhimg.so) through its drivers."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "himg_b200", "_lib")


def _run(args):
    return subprocess.run(args, capture_output=True, text=True, timeout=300)


def test_chimg_dhimg_round_trip(tmp_path, port):
    from himg_b200 import build as hb

    hb.build_host()
    himg = str(tmp_path / "a.himg")
    ppm = str(tmp_path / "a.ppm")
    r = _run([os.path.join(LIBDIR, "chimg"), "-q", "90", "synthetic:256x136x3:5:6", himg])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Low resolution data:" in r.stdout and "Full resolution data:" in r.stdout and "Compressed size:" in r.stdout
    img = port.synth(256, 136, 3, 5, 6)
    want = port.encode(img, 90, True)
    assert open(himg, "rb").read() == want
    r = _run([os.path.join(LIBDIR, "dhimg"), himg, ppm])
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(ppm, "rb").read()
    assert raw.startswith(b"P6\n256 136\n255\n")
    got = np.frombuffer(raw[len(b"P6\n256 136\n255\n"):], np.uint8).reshape(136, 256, 3)
    assert np.array_equal(got, port.decode(want))
    # PPM in -> same bitstream; -rgb switches the colour space off
    r = _run([os.path.join(LIBDIR, "chimg"), "-rgb", ppm, himg])
    assert r.returncode == 0
    assert open(himg, "rb").read() == port.encode(got, 50, False)


def test_benchmark_driver(tmp_path, port):
    himg = str(tmp_path / "b.himg")
    open(himg, "wb").write(port.encode(port.synth(512, 512, 3, 1, 6), 50, True))
    r = _run([os.path.join(LIBDIR, "benchmark"), himg])
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Iteration 30/30" in r.stdout and "Average:" in r.stdout and "Min:" in r.stdout
    r = _run([os.path.join(LIBDIR, "benchmark"), "-e", "synthetic:512x512x3"])
    assert r.returncode == 0 and "Average:" in r.stdout


def test_dhimg_rejects_like_the_reference(tmp_path, port):
    himg = str(tmp_path / "c.himg")
    open(himg, "wb").write(port.encode(port.synth(512, 512, 3, 7, 6), 20, True))  # reference decoder refuses this
    r = _run([os.path.join(LIBDIR, "dhimg"), himg, str(tmp_path / "c.ppm")])
    assert r.returncode != 0 and "Unable to decode image." in r.stdout
    env = dict(os.environ, HIMG_LENIENT="1")
    r = subprocess.run([os.path.join(LIBDIR, "dhimg"), himg, str(tmp_path / "c.ppm")], capture_output=True, text=True, env=env)
    assert r.returncode == 0
