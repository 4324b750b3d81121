import sys, torch
sys.path.insert(0, '/root/repo')
from bruno_b200 import _lib, slab
N, H, W = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (640, 28, 28)
g = torch.Generator().manual_seed(0)
V = torch.randn(1, 3, 3, 64, 64, generator=g) * 0.05
wf = torch.empty(1, _lib.query('bruno_conv_wpack_bytes'), dtype=torch.uint8, device='cuda')
_lib.call('bruno_conv_wn_pack', V.cuda(), torch.ones(1, 64, device='cuda'), wf, None, 1)
sl = slab.alloc_slabs(3, N, H, W, 'cuda')
sl[0].random_(0, 255)   # garbage activations are fine for timing
bias = torch.zeros(64, device='cuda')
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 1
_lib.load().bruno_debug_conv_variant(variant)
for mode in (0, 3):
    for _ in range(3):
        _lib.call('bruno_conv3x3', mode, sl[0], wf[0], bias if mode < 2 else None, 0, sl[2] if mode else None, sl[2] if mode == 3 else None, sl[1], N, H, W)
    tr = torch.zeros(64 * 16, dtype=torch.int64, device='cuda')
    _lib.load().bruno_debug_conv_trace(tr.data_ptr())
    _lib.call('bruno_conv3x3', mode, sl[0], wf[0], bias if mode < 2 else None, 0, sl[2] if mode else None, sl[2] if mode == 3 else None, sl[1], N, H, W)
    torch.cuda.synchronize()
    _lib.load().bruno_debug_conv_trace(None)
    t = tr.view(64, 16).cpu()
    t0 = int(t[0, 0])
    print('mode', mode, ' columns: load_issued slab_landed mma_issued acc_ready epi_done store_issued store_read_done (cycles since first load)')
    for i in range(30):
        print(i, 'load %d landed %d mma_issued %d | epi start %d tmem_free %d done %d | store %d' % tuple(int(t[i, k] - t0) for k in (0, 1, 2, 3, 6, 4, 5)))
