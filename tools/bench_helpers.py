"""Timing of the memory-bound helpers of a conv coupling (c1 forward, heads + coupling forward) at the bn2_omniglot_tp shapes.
Usage: python tools/bench_helpers.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib, slab


def timed(fn, rep=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(rep): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / rep * 1e3


for (N, H, W, C) in ((640, 28, 28, 1), (640, 14, 14, 4), (640, 14, 14, 2)):
    g = torch.Generator().manual_seed(0)
    x = torch.rand(N, H, W, C, generator=g).cuda()
    hs = slab.nhwc_to_slab(torch.randn(N, H, W, 64, generator=g).cuda())
    out = slab.alloc_slabs(1, N, H, W, 'cuda')[0]
    V1 = (torch.randn(C, 64, generator=g) * 0.3).cuda(); g1 = torch.ones(64).cuda(); b1 = torch.zeros(64).cuda()
    Ks = (torch.randn(64, C, generator=g) * 0.05).cuda(); Kt = (torch.randn(64, C, generator=g) * 0.05).cuda()
    bs = torch.zeros(C).cuda(); bt = torch.zeros(C).cuda()
    ld = torch.zeros(N).cuda(); ld2 = torch.empty(N).cuda(); y = torch.empty_like(x)
    t_c1 = timed(lambda: _lib.call('bruno_c1_fwd', x, V1, g1, b1, 0, out, 0, N, H, W, C, 0, 1))
    t_hd = timed(lambda: _lib.call('bruno_heads_coupling', x, hs, Ks, bs, Kt, bt, None, None, ld, y, ld2, 0, N, H, W, C, 0, 0))
    mb = hs.numel() / 1e6
    print('%dx%dx%d C=%d: c1 forward %.1f us (%.2f TB/s written), heads + coupling forward %.1f us (%.2f TB/s read)' % (
        N, H, W, C, t_c1, mb / t_c1, t_hd, mb / t_hd))
