"""Debug helper: REP back-to-back launches of bruno_conv3x3 (mode, N, H, W from argv) as a dependent ping-pong chain,
line-tile kernel vs per-tap kernel (bit-exact comparison of the final slab)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib, slab
mode, N, H, W = (int(a) for a in sys.argv[1:5])
REP = int(sys.argv[5]) if len(sys.argv) > 5 else 20
g = torch.Generator().manual_seed(0)
V = torch.randn(1, 3, 3, 64, 64, generator=g) * 0.02
wb = _lib.query('bruno_conv_wpack_bytes')
wf = torch.empty(1, wb, dtype=torch.uint8, device='cuda'); wbw = torch.empty(1, wb, dtype=torch.uint8, device='cuda')
_lib.call('bruno_conv_wn_pack', V.cuda(), torch.ones(1, 64, device='cuda'), wf, wbw, 1)
x0, a1, a2 = [slab.nhwc_to_slab(torch.randn(N, H, W, 64, generator=g).cuda() * 0.5) for _ in range(3)]
bias = torch.randn(64, generator=g).cuda() * 0.1
lib = _lib.load()
res = []
for v in (1, 2):
    lib.bruno_debug_conv_variant(v)
    sl = slab.alloc_slabs(2, N, H, W, 'cuda')
    sl[0].copy_(x0)
    PP = int(os.environ.get('PINGPONG', '1'))
    for i in range(REP):
        if PP == 0: i = 0
        if PP == 2: i = 1
        _lib.call('bruno_conv3x3', mode, sl[i & 1], wf[0] if mode in (0, 1, 4) else wbw[0], bias, 0, a1, a2, sl[1 - (i & 1)], N, H, W)
    torch.cuda.synchronize()
    print('variant', v, 'ok', flush=True)
    res.append(sl[REP & 1].clone())
print('bitwise equal:', bool(torch.equal(res[0], res[1])))
