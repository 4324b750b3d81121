import sys, importlib, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from oracle import bruno_oracle as O, bruno_oracle_cond as OC
import test_conditional_gpu as T
for seed in (3, 7):
    for nc in (1, 0):
        spec, P, x, yl, cfg, xg, ylg = T.setup(2, 3, seed=seed, n_context=nc)
        with torch.no_grad():
            lp_o, _ = OC.build_model(spec, P, x, yl, n_context=nc)
            lp_i = cfg.build_model(xg, ylg)[0]
        cfg.set_trainable(True)
        lp_t = cfg.build_model(xg, ylg)[0]
        cfg.set_trainable(False)
        rel = lambda a, b: ((a.double().cpu() - b).abs() / b.abs().clamp_min(1.0)).max().item()
        print('seed %d n_context %d: inference vs oracle %.3e   training-graph vs oracle %.3e   |lp| ~ %.3e' % (seed, nc, rel(lp_i, lp_o), rel(lp_t.detach(), lp_o), lp_o.abs().mean().item()))
