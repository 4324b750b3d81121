"""Pipeline timeline of CTA 0 of the forward/dgrad conv kernel (bruno_debug_conv_trace; DEBUG build only):
kernel entry, set-up, previous-grid-complete, weights landed, then per input tile: load issued, issuer past its waits,
MMAs issued, and per output tile: accumulator ready, epilogue done, store issued; and the kernel's end.  Cycles of SM 0's clock64().
Usage: python tools/trace_conv.py [N H W]      (default 640 14 14: the launch-overhead-bound layer shape)"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib, slab
N, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (640, 14, 14)
g = torch.Generator().manual_seed(0)
V = torch.randn(1, 3, 3, 64, 64, generator=g) * 0.05
wf = torch.empty(1, _lib.query('bruno_conv_wpack_bytes'), dtype=torch.uint8, device='cuda')
_lib.call('bruno_conv_wn_pack', V.cuda(), torch.ones(1, 64, device='cuda'), wf, None, 1)
sl = slab.alloc_slabs(3, N, H, W, 'cuda')
sl[0].random_(0, 255)
bias = torch.zeros(64, device='cuda')
lib = _lib.load()
if os.environ.get('BRUNO_CONV_VARIANT'): lib.bruno_debug_conv_variant(int(os.environ['BRUNO_CONV_VARIANT']))   # 1 per-tap, 2 line-tile


def launch(mode):
    _lib.call('bruno_conv3x3', mode, sl[0], wf[0], bias if mode < 2 else None, 0, sl[2] if mode else None,
              sl[2] if mode == 3 else None, sl[1], N, H, W)


for mode in (0, 1, 2, 3):
    for _ in range(3):
        launch(mode)
    tr = torch.zeros(64 * 16, dtype=torch.int64, device='cuda')
    lib.bruno_debug_conv_trace(tr.data_ptr())
    launch(mode)                       # back to back with a predecessor, like inside a step
    launch(mode)
    torch.cuda.synchronize()
    lib.bruno_debug_conv_trace(None)
    t = tr.view(64, 16).cpu()
    t0 = int(t[0, 8])
    rel = lambda v: int(v) - t0
    print('mode %d  %dx%dx%d   (cycles since kernel entry of CTA 0; second of two back-to-back launches)' % (mode, N, H, W))
    print('  setup done %d | previous grid complete %d | weights landed %d | kernel end %d' % (
        rel(t[0, 9]), rel(t[0, 10]), rel(t[0, 11]), rel(t[0, 12])))
    if int(t[62, 0]):
        print('  prologue: barriers initialised %d | stage rows zeroed %d | TMEM allocated %d' % tuple(rel(t[62, k]) for k in (0, 1, 2)))
    cyc, ns = int(t[0, 12] - t[0, 8]), int(t[0, 14] - t[0, 13])
    print('  CTA 0 lifetime: %d cycles in %d ns = %.0f MHz effective SM clock' % (cyc, ns, 1e3 * cyc / max(ns, 1)))
    for i in range(int(os.environ.get('TRACE_TILES', '32'))):
        if int(t[i, 0]) == 0:
            break
        print('  tile %2d: load issued %6d | mma start %6d issued %6d | acc ready %6d epi done %6d | store issued %6d' % (
            (i,) + tuple(rel(t[i, k]) for k in (0, 1, 2, 3, 4, 5))))
