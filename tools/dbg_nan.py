import sys, importlib, torch, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import bruno_oracle as O
name = 'an2_cifar_tp'
spec = O.spec_for(name)
P = O.init_params(spec, seed=5, dtype=torch.float64, perturb_process=True)
x = O.synthetic_batch(spec, 1, 2, seed=6, kind='bytes').float().double()
O.data_dependent_init(spec, P, x)
P = O.cast_params(O.cast_params(P, torch.float32), torch.float64)
cfg = importlib.import_module('bruno_b200.config_rnn.' + name)
xg = x.float().cuda()
cfg.build_model(xg[:1]); cfg.model.load_tf_variables(P)
lp, lat, pri, z = cfg.build_model(xg)
print('z absmax', z.abs().max().item(), 'lp', lp)
sl = cfg.student_layer
zd = z.detach().clone().requires_grad_(True)
l2, p2 = sl.sequence_log_likelihoods(zd)
g = torch.autograd.grad(l2.sum(), [zd, sl.var_vbl, sl.nu_vbl, sl.corr_vbl])
for n, t in zip(('dz', 'dvar', 'dnu', 'dcorr'), g):
    print(n, 'nan count', torch.isnan(t).sum().item(), 'absmax', t[~torch.isnan(t)].abs().max().item())
bad = torch.isnan(g[1]).nonzero()
print('bad dims', bad[:10].tolist())
if len(bad):
    d = bad[0][1].item()
    print('z at bad dim', zd[0, :, d], 'var_vbl', sl.var_vbl[0, d].item(), 'nu_vbl', sl.nu_vbl[0, d].item(), 'corr_vbl', sl.corr_vbl[0, d].item())
    from bruno_b200.process import build_table
    tab = build_table(0, sl.var_vbl.detach(), sl.nu_vbl.detach(), sl.corr_vbl.detach(), sl.mu, None, 2.0, 0, 2, True)
    D = 3072
    t = tab.view(-1, D)
    print('table rows at bad dim', t[:, d].tolist())
