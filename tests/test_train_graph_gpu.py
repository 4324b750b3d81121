"""GPU: the CUDA-graphed training step (Trainer.enable_cuda_graph: forward + hand-written backward replayed as one graph
launch, allreduce + optimiser eager) follows exactly the same trajectory as eager launches."""
import importlib

import pytest
import torch

from oracle import bruno_oracle as O

pytestmark = pytest.mark.gpu


def _trainer(name, P, xs):
    from bruno_b200.trainer import Trainer
    cfg = importlib.reload(importlib.import_module('bruno_b200.config_rnn.' + name))
    cfg.optimizer = 'rmsprop'
    tr = Trainer(cfg, 'cuda:0')
    tr.materialize(xs[0], init=False)
    cfg.model.load_tf_variables(P)
    return tr


@pytest.mark.parametrize('name,B,T', [('en2_noconv_mnist_tp', 4, 5), ('bn2_omniglot_tp', 2, 3)])
def test_graphed_train_step_equals_eager(name, B, T):
    spec = O.spec_for(name)
    P = O.init_params(spec, seed=2, dtype=torch.float64, perturb_process=True)
    xs = [O.synthetic_batch(spec, B, T, seed=30 + i).float().cuda() for i in range(4)]
    O.data_dependent_init(spec, P, xs[0].double().cpu())
    P = O.cast_params(P, torch.float32)
    losses = {}
    flats = {}
    for mode in ('eager', 'graph'):
        tr = _trainer(name, P, xs)
        if mode == 'graph':
            tr.enable_cuda_graph(xs[0])
        ls = []
        for step in range(4):
            lr = 1e-3 * 0.9 ** step                       # the learning rate changes every iteration (train.py:165-166)
            ls.append(tr.train_step(xs[step], lr, 0.1 * step).item())
        losses[mode], flats[mode] = ls, tr.flat.detach().clone()
    print('\n[%s] losses eager %s | graph %s' % (name, losses['eager'], losses['graph']))
    if name.startswith('en2'):
        # dense path: deterministic kernels (ordered split-K fix-up) -> bit-identical
        assert losses['eager'] == losses['graph']
        assert torch.equal(flats['eager'], flats['graph'])
    else:
        # the conv wgrad accumulates with red.global.add: run-to-run differences of a few ulp, graph or not
        # (measured: losses agree to 3e-5, parameters after four RMSProp steps -- sign-like updates of lr per element,
        # which turn a last-bit gradient difference near zero into +-lr -- to 7e-5 relative L2)
        for a, b in zip(losses['eager'], losses['graph']):
            assert abs(a - b) <= 2e-4 * abs(a)
        assert ((flats['eager'] - flats['graph']).norm() / flats['eager'].norm()).item() <= 5e-4


def test_graph_falls_back_to_eager_for_other_shapes():
    name = 'en2_noconv_mnist_tp'
    spec = O.spec_for(name)
    P = O.cast_params(O.init_params(spec, seed=2, dtype=torch.float64), torch.float32)
    xs = [O.synthetic_batch(spec, 4, 5, seed=40).float().cuda(), O.synthetic_batch(spec, 2, 5, seed=41).float().cuda()]
    tr = _trainer(name, P, xs)
    tr.enable_cuda_graph(xs[0])
    a = tr.train_step(xs[0], 1e-3, 0.1).item()
    b = tr.train_step(xs[1], 1e-3, 0.1).item()            # different batch size: not the captured shape
    assert a == a and b == b
