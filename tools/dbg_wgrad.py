import sys, torch
sys.path.insert(0, '/root/repo')
from bruno_b200 import _lib, slab
def run(L, N, H, W, batched):
    xs = slab.alloc_slabs(L, N, H, W, 'cuda'); gs = slab.alloc_slabs(L, N, H, W, 'cuda')
    g = torch.Generator().manual_seed(1)
    for l in range(L):
        xs[l].copy_(slab.nhwc_to_slab(torch.randn(N, H, W, 64, generator=g).cuda()))
        gs[l].copy_(slab.nhwc_to_slab(torch.randn(N, H, W, 64, generator=g).cuda()))
    dW = torch.zeros(L, 3, 3, 64, 64, device='cuda'); db = torch.zeros(L, 64, device='cuda')
    try:
        if batched: _lib.call('bruno_conv3x3_wgrad_batched', xs, gs, dW, db, L, N, H, W)
        else: _lib.call('bruno_conv3x3_wgrad', xs[0], gs[0], dW[0], db[0], N, H, W)
        torch.cuda.synchronize()
        print('ok', L, N, H, W, batched, float(dW.abs().sum()))
    except Exception as e:
        print('FAIL', L, N, H, W, batched, str(e)[:100]); sys.exit(1)
for args in [(1, 3, 28, 28, False), (1, 6, 28, 28, False), (1, 6, 28, 28, True), (2, 6, 28, 28, True), (4, 6, 28, 28, True), (16, 6, 14, 14, True)]:
    run(*args)
