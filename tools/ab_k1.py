#!/usr/bin/env python
"""tools/ab_k1.py -- channel kernel (ccgpu_awgn_llr) write rate for a few n; CCGPU_LIB selects the library under test."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import channelcoding_b200 as cc

ctx = cc.Context(0)
ctx.use_torch_stream()
for n, frames in ((63, 1 << 24), (15, 1 << 26), (127, 1 << 23), (255, 1 << 22)):
    y = torch.empty((frames, n), dtype=torch.float32, device="cuda")
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.awgn_llr(n, np.float32(0.6), 0, 1, 0, frames, out=y)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print("%s n=%d: %.3f ms  %.3e frames/s  %.1f GB/s  checksum %.6f" % (os.environ.get("CCGPU_LIB", "libccgpu.so").split("/")[-1], n, best,
          frames / best * 1e3, frames * n * 4 / best / 1e6, float(y[:1000].double().sum())), flush=True)
    del y
