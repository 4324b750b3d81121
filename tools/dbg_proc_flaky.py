"""Debug: the running-state (STATE) instantiation of the forward ring scan, two chunks per repetition with fresh buffers:
what must precede chunk 1 for one of its rows to go wrong?  B = 1030, T = 10 + 10, D = 784, GP."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from bruno_b200 import _lib
from bruno_b200.process import build_table
import test_process_gpu as T
B, TT, D = 1030, 20, 784
kind = 'gp'
layer, P = T._layer(kind, D, seed=3, mu=0.)
g = torch.Generator().manual_seed(B + D)
z = (torch.randn(B, TT, D, generator=g) * 1.3 + 0.5 * torch.randn(B, 1, D, generator=g)).float().cuda()
_lib.query('bruno_process_set_fwd_variant', 9)
ref = layer.sequence_log_likelihoods(z)[0].detach().clone()
T1 = TT // 2
for mode in ('ring, ring', 'ring, staged', 'ring, ring + sync before chunk 1', 'ring, ring, nan-prefilled lp', 'ring, ring, zs fixed', 'ring, ring, lps fixed'):
    bad1 = 0
    zs_f = [z[:, :T1].contiguous(), z[:, T1:].contiguous()]
    lps_f = [torch.empty(B, T1, device='cuda') for _ in range(2)]
    for rep in range(100):
        s1 = torch.zeros(B, D, device='cuda'); s2 = torch.zeros(B, D, device='cuda')
        zs = zs_f if 'zs fixed' in mode else [z[:, :T1].contiguous(), z[:, T1:].contiguous()]
        tabs = [build_table(layer.kind, layer.var_vbl.detach(), None, layer.corr_vbl.detach(), layer.mu, None, layer.min_nu, n0, T1, False) for n0 in (0, T1)]
        lps = lps_f if 'lps fixed' in mode else [torch.full((B, T1), float('nan'), device='cuda') if 'nan' in mode else torch.empty(B, T1, device='cuda') for _ in range(2)]
        scr = torch.empty(1, device='cuda')
        if 'sync' in mode: torch.cuda.synchronize()
        for c in range(2):
            _lib.query('bruno_process_set_fwd_variant', 9 if (c == 1 and 'staged' in mode) else 16)
            _lib.call('bruno_process_fwd', layer.kind, zs[c], layer.mu, tabs[c], s1, s2, lps[c], None, scr, B, T1, D)
            if c == 0:
                d = (lps[0] - ref[:, :T1]).abs()
                if bool((~(d <= 0.01)).any()): bad1 += 1
    print('%-36s: chunk 1 wrong in %d of 100' % (mode, bad1))
_lib.query('bruno_process_set_fwd_variant', 0)
