"""Training / NLL step time of Conditional BRUNO (config_conditional/m1_shapenet, B=4, T=16, 32x32x1, n_context=1)."""
import sys, importlib, json, torch
sys.path.insert(0, '/root/repo')
from bruno_b200.config_conditional import defaults
defaults.seq_len, defaults.n_context = 16, 1
cfg = importlib.import_module('bruno_b200.config_conditional.m1_shapenet')
from bruno_b200.config_conditional.train import ConditionalTrainer
dev = torch.device('cuda', 0)
it = cfg.train_data_iter.generate()
batches = [next(it) for _ in range(4)]
xs = [torch.from_numpy(b[0]).to(dev) for b in batches]
ys = [torch.from_numpy(b[1]).to(dev) for b in batches]
tr = ConditionalTrainer(cfg, dev).materialize(xs[0], ys[0], init=True)
def timed(fn, k=10):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
ms_train = timed(lambda i: tr.train_step(xs[i % 4], ys[i % 4], 1e-3, 0.5))
ms_eval = timed(lambda i: tr.eval_step(xs[i % 4], ys[i % 4]))
print(json.dumps({'metric': 'm1_shapenet_train_seqs_per_sec', 'value': cfg.batch_size / (ms_train * 1e-3), 'ms_per_step': ms_train,
                  'eval_nll_ms': ms_eval, 'eval_nll_seqs_per_sec': cfg.batch_size / (ms_eval * 1e-3), 'params': tr.n_params,
                  'config': {'workload': 'Conditional BRUNO m1_shapenet: 14 label-conditioned conv couplings (R=6) + 6 dense, GP, n_context 1',
                             'batch': cfg.batch_size, 'seq_len': cfg.seq_len, 'obs_shape': list(cfg.obs_shape[1:])}}))
