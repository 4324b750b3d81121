"""Timing of the memory-bound helpers of a conv coupling (c1 forward / backward, heads + coupling forward / backward) at the
bn2_omniglot_tp shapes.
Usage: python tools/bench_helpers.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib, slab


def timed(fn, rep=20):
    """us per launch, GPU side: `rep` launches replayed from one CUDA graph (eager launches through ctypes are host-bound
    at ~15 us per call, more than most of these kernels take)"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(rep): fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * rep) * 1e3


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
    gs = slab.alloc_slabs(1, N, H, W, 'cuda')[0]
    dy = torch.randn(N, H, W, C, generator=g).cuda(); dld = torch.randn(N, generator=g).cuda(); dx = torch.zeros_like(x)
    dKs, dKt, dbs, dbt = torch.empty_like(Ks), torch.empty_like(Kt), torch.empty_like(bs), torch.empty_like(bt)
    scratch = torch.empty(_lib.query('bruno_heads_bwd_scratch_floats', C), device='cuda')
    dW1, db1 = torch.zeros_like(V1), torch.zeros(64).cuda()
    t_hb = timed(lambda: _lib.call('bruno_heads_coupling_bwd', x, hs, Ks, bs, Kt, bt, dy, dld, dx, gs, dKs, dbs, dKt, dbt, scratch,
                                   N, H, W, C, 0))
    t_cb = timed(lambda: _lib.call('bruno_c1_bwd', x, V1, g1, hs, dx, dW1, db1, N, H, W, C, 0, 1))
    mb = hs.numel() / 1e6
    print('%dx%dx%d C=%d: heads + coupling backward %.1f us (%.2f TB/s read + written), c1 backward %.1f us (%.2f TB/s read)' % (
        N, H, W, C, t_hb, 2 * mb / t_hb, t_cb, mb / t_cb))
    print('%dx%dx%d C=%d: c1 forward %.1f us (%.2f TB/s written), heads + coupling forward %.1f us (%.2f TB/s read)' % (
        N, H, W, C, t_c1, mb / t_c1, t_hd, mb / t_hd))
