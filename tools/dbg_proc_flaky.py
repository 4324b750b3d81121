import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from bruno_b200 import _lib
from bruno_b200.process import build_table
import test_process_gpu as T
kind, B, TT, D = 'gp', 1030, 20, 784
layer, P = T._layer(kind, D, seed=3, mu=0.)
g = torch.Generator().manual_seed(B + D)
z = (torch.randn(B, TT, D, generator=g) * 1.3 + 0.5 * torch.randn(B, 1, D, generator=g)).float().cuda()
_lib.query('bruno_process_set_fwd_variant', 9)
ref = layer.sequence_log_likelihoods(z)[0].detach().clone()
T1 = TT // 2
bad = {9: 0, 0: 0}
for rep in range(60):
    for v in (9, 0):
        _lib.query('bruno_process_set_fwd_variant', v)
        s1 = torch.zeros(B, D, device='cuda'); s2 = torch.zeros(B, D, device='cuda')
        lps = []
        for n0, zz in ((0, z[:, :T1].contiguous()), (T1, z[:, T1:].contiguous())):
            Tn = zz.shape[1]
            tab = build_table(layer.kind, layer.var_vbl.detach(), None, layer.corr_vbl.detach(), layer.mu, None, layer.min_nu, n0, Tn, False)
            lp = torch.empty(B, Tn, device='cuda')
            scratch = torch.empty(max(_lib.query('bruno_process_scratch_floats', B, Tn, D), 1), device='cuda')
            _lib.call('bruno_process_fwd', layer.kind, zz, layer.mu, tab, s1, s2, lp, None, scratch, B, Tn, D)
            lps.append(lp)
        lp_cat = torch.cat(lps, 1)
        d = (lp_cat - ref).abs()
        if d.max().item() > 0.01:
            bad[v] += 1
            idx = (d > 0.01).nonzero()
            print('rep', rep, 'variant', v, 'max', d.max().item(), 'count', idx.shape[0], 'rows', sorted(set(idx[:, 0].tolist()))[:12], 'cols', sorted(set(idx[:, 1].tolist())))
print('bad counts', bad)
_lib.query('bruno_process_set_fwd_variant', 0)
# non-state ring kernel (the product path): same shape, 200 repetitions
bad_ns = 0
for rep in range(200):
    lp = layer.sequence_log_likelihoods(z)[0]
    if (lp - ref).abs().max().item() > 0.01:
        bad_ns += 1
print('non-state ring kernel: bad', bad_ns, 'of 200')
