"""Per-mode timing of the forward/dgrad conv kernel on one layer shape (default: the bn2_omniglot_tp 28x28 layer)."""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib, slab
N, H, W = 640, 28, 28
if len(sys.argv) > 1: N, H, W = (int(a) for a in sys.argv[1:4])
g = torch.Generator().manual_seed(0)
V = torch.randn(1, 3, 3, 64, 64, generator=g) * 0.05
wf = torch.empty(1, _lib.query('bruno_conv_wpack_bytes'), dtype=torch.uint8, device='cuda')
_lib.call('bruno_conv_wn_pack', V.cuda(), torch.ones(1, 64, device='cuda'), wf, None, 1)
sl = slab.alloc_slabs(4, N, H, W, 'cuda')
bias = torch.zeros(64, device='cuda')
flops = 2.0 * N * H * W * 9 * 64 * 64
for variant in (1,):
    for mode in (0, 1, 2, 3):
        def run():
            _lib.call('bruno_conv3x3', mode, sl[0], wf[0], bias if mode in (0, 1) else None, 0, sl[2] if mode else None,
                      sl[3] if mode == 3 else None, sl[1], N, H, W)
        for _ in range(5): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(40): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 40 * 1e3
        print('mode %d: %.1f us  %.0f TFLOP/s' % (mode, us, flops / us * 1e-6))
