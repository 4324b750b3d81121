"""Timing of the batched wgrad launch of one conv coupling (L = 16 layers) at the two bn2_omniglot_tp shapes.
Usage: python tools/bench_wgrad.py [variant]   (bruno_debug_conv_variant: 0 auto / line-tile, 1 per-tap)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib, slab
lib = _lib.load()
if len(sys.argv) > 1: lib.bruno_debug_conv_variant(int(sys.argv[1]))
L = 16
shapes = [tuple(int(v) for v in a.split(',')) for a in sys.argv[2:]] or [(640, 28, 28), (640, 14, 14)]
for (N, H, W) in shapes:
    xs = slab.alloc_slabs(L, N, H, W, 'cuda'); gs = slab.alloc_slabs(L, N, H, W, 'cuda')
    xs.random_(0, 2); gs.random_(0, 2)
    dW = torch.zeros(L, 3, 3, 64, 64, device='cuda'); db = torch.zeros(L, 64, device='cuda')
    f = lambda: _lib.call('bruno_conv3x3_wgrad_batched', xs, gs, dW, db, L, N, H, W)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 5 * 1e3
    flops = 2.0 * L * N * H * W * 9 * 64 * 64
    print('%dx%dx%d L=%d: %.1f us per launch  %.0f TFLOP/s  (%.2f TB/s of slab reads)' % (N, H, W, L, us, flops / us * 1e-6, 2 * xs.numel() / us * 1e-6))
    del xs, gs
