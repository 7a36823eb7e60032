import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import himg_b200
from himg_b200.synth import synth_images
W, H, N, B = 1920, 1080, 3, 128
ctx = himg_b200.Context(0)
px = synth_images(B, W, H, N, 1, 6)
h_px = torch.empty((B, H, W, N), dtype=torch.uint8, pin_memory=True); h_px.copy_(px)
bound = himg_b200.encode_bound(W, H, N)
h_out = torch.empty((B * bound,), dtype=torch.uint8, pin_memory=True)
off = np.zeros(B + 1, np.uint64); sz = np.zeros(B, np.uint32)
for _ in range(2): ctx.encode_batch_host(h_px, 50, True, out=h_out, offsets=off, sizes=sz)
os.environ["HIMG_DEBUG_PIPE"] = "1"
ctx.encode_batch_host(h_px, 50, True, out=h_out, offsets=off, sizes=sz)
