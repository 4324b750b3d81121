"""Debug: how often does the running-state (STATE) instantiation of a forward scan variant return a wrong row?
Chunk 1 only (s1 = 0), B = 1030, T = 10, D = 784, GP and TP; variants 9 (register-staged), 16 (ring, CPF=2 = default shape),
10 (ring, CPF=1), 15 (ring, CPF=0)."""
import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from bruno_b200 import _lib
from bruno_b200.process import build_table
import test_process_gpu as T
B, TT, D = 1030, 20, 784
for kind in ('gp', 'tp'):
    layer, P = T._layer(kind, D, seed=3, mu=0.)
    g = torch.Generator().manual_seed(B + D)
    z = (torch.randn(B, TT, D, generator=g) * 1.3 + 0.5 * torch.randn(B, 1, D, generator=g)).float().cuda()
    _lib.query('bruno_process_set_fwd_variant', 9)
    ref = layer.sequence_log_likelihoods(z)[0].detach().clone()
    T1 = TT // 2
    zz = z[:, :T1].contiguous()
    nu = layer.nu_vbl.detach() if kind == 'tp' else None
    tab = build_table(layer.kind, layer.var_vbl.detach(), nu, layer.corr_vbl.detach(), layer.mu, None, layer.min_nu, 0, T1, False)
    scratch = torch.empty(max(_lib.query('bruno_process_scratch_floats', B, T1, D), 1), device='cuda')
    for v in (9, 16, 10, 15):
        _lib.query('bruno_process_set_fwd_variant', v)
        bad = 0
        rows = set()
        for rep in range(100):
            s1 = torch.zeros(B, D, device='cuda'); s2 = torch.zeros(B, D, device='cuda')
            lp = torch.empty(B, T1, device='cuda')
            _lib.call('bruno_process_fwd', layer.kind, zz, layer.mu, tab, s1, s2, lp, None, scratch, B, T1, D)
            d = (lp - ref[:, :T1]).abs()
            if d.max().item() > 0.01:
                bad += 1
                rows |= set((d > 0.01).nonzero()[:, 0].tolist())
        print(kind, 'variant', v, 'bad', bad, 'of 100', sorted(rows)[:10])
_lib.query('bruno_process_set_fwd_variant', 0)
