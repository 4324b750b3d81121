import sys, importlib, torch
sys.path.insert(0, '/root/repo')
from oracle import bruno_oracle as O, bruno_oracle_cond as OC
B, T = 2, 3
spec = OC.CondSpec()
P = OC.init_params(spec, seed=11, dtype=torch.float64)
x, yl = OC.synthetic_batch(spec, B, T, seed=12)
x, yl = x.float().double(), yl.float().double()
P = O.cast_params(O.cast_params(P, torch.float32), torch.float64)
from bruno_b200.config_conditional import defaults
defaults.n_context, defaults.seq_len = 1, T
cfg = importlib.import_module('bruno_b200.config_conditional.m1_shapenet')
xg, ylg = x.float().cuda(), yl.float().cuda()
cfg.build_model(xg[:1], ylg[:1]); cfg.load_tf_variables(P)
cfg.build_model(xg, ylg, init=True)
OC.data_dependent_init(spec, P, x, yl, n_context=1)
rows = []
for name, view in cfg.named_tf_variables():
    if not (name.endswith('/g') or name.endswith('/b')): continue
    ref = P[name].double(); got = view.detach().double().cpu().reshape(ref.shape)
    rows.append(((got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-6), name, ref.abs().max().item()))
import collections
by = collections.OrderedDict()
for e, n, m in rows:
    layer = n.split('/')[1]
    by.setdefault(layer, []).append((e, n, m))
for layer, v in by.items():
    w = max(v)
    print('%-18s n=%3d worst %.2e at %s (|ref| %.2e)' % (layer, len(v), w[0], w[1].split('/', 2)[2], w[2]))
