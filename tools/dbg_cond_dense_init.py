import sys, torch
sys.path.insert(0, '/root/repo')
from oracle import bruno_oracle as O, bruno_oracle_cond as OC
from bruno_b200 import nn_extra_nvp_conditional as C, ddinit
torch.manual_seed(0)
N, D, L = 6, 1024, 32
layer = C.CouplingLayerDense('even', name='even_0', n_units=256)
x = torch.randn(N, 4, 4, 64, device='cuda')
yl = torch.randn(N, L, device='cuda')
ld = torch.zeros(N, device='cuda')
layer._build(D, L, x.device)
for k, v in layer.P.items():
    if k.endswith('kernel'): v.normal_(0, 0.05)
P = {n: v.detach().double().cpu().clone() for n, v in layer.named_tf_variables(prefix='')}
out, ldo = ddinit.cond_dense_coupling_init(layer, x, ld, yl)
idx = torch.arange(D) % 2
mask = idx.double().view(1, 4, 4, 64).expand(N, 4, 4, 64) if False else (torch.arange(D) % 2).double().view(4, 4, 64).expand(N, 4, 4, 64)
xo = x.double().cpu()
s, t = OC.s_t_dense(xo * mask, mask, yl.double().cpu(), P, 'even_0', True)
for n, v in layer.named_tf_variables(prefix=''):
    if n.endswith('/g') or n.endswith('/b'):
        ref = P[n]; got = v.detach().double().cpu()
        print('%-50s rel err %.2e' % (n, (got - ref).abs().max().item() / ref.abs().max().item()))
