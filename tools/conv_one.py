"""Debug helper: one bruno_conv3x3 launch (mode, N, H, W from argv) checked against the per-tap kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib, slab
mode, N, H, W = (int(a) for a in sys.argv[1:5])
g = torch.Generator().manual_seed(0)
V = torch.randn(1, 3, 3, 64, 64, generator=g) * 0.05
wb = _lib.query('bruno_conv_wpack_bytes')
wf = torch.empty(1, wb, dtype=torch.uint8, device='cuda'); wbw = torch.empty(1, wb, dtype=torch.uint8, device='cuda')
_lib.call('bruno_conv_wn_pack', V.cuda(), torch.ones(1, 64, device='cuda'), wf, wbw, 1)
xs, a1, a2 = [slab.nhwc_to_slab(torch.randn(N, H, W, 64, generator=g).cuda()) for _ in range(3)]
bias = torch.randn(64, generator=g).cuda()
out = slab.alloc_slabs(2, N, H, W, 'cuda')
lib = _lib.load()
for v in (1, 2):
    lib.bruno_debug_conv_variant(v)
    _lib.call('bruno_conv3x3', mode, xs, wf[0] if mode in (0, 1, 4) else wbw[0], bias, 0, a1, a2, out[v - 1], N, H, W)
    torch.cuda.synchronize()
    print('variant', v, 'ok', flush=True)
a, b = slab.slab_to_nhwc(out[0], N, H, W, check_padding=True), slab.slab_to_nhwc(out[1], N, H, W, check_padding=True)
print('max abs diff line vs tap:', float((a - b).abs().max()), 'bitwise equal:', bool(torch.equal(out[0], out[1])), 'ref max', float(b.abs().max()))
