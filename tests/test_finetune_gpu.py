"""GPU: episodic fine-tuning -- the softmax-over-candidates loss kernel against numbers the
REFERENCE's own loss function produced (tests/golden/ref_finetune_loss.npz, config_rnn/bn2_omniglot_tp_ft_1s_20w.py:
222-226 executed over tests/tf_shim), the config surface, and the train -> train_finetune -> prior-samples drivers end to
end."""
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_episode_loss_kernel_matches_reference_loss(golden_dir):
    from bruno_b200 import losses
    g = np.load(os.path.join(golden_dir, 'ref_finetune_loss.npz'))
    M, B, T = int(g['meta_batch_size']), int(g['batch_size']), int(g['seq_len'])
    lp = torch.from_numpy(g['log_probs']).float().cuda().requires_grad_(True)
    loss = losses.episode_softmax_loss(lp, M, B, T)
    loss.backward()
    ref_loss = float(g['loss'])
    # the fixture's log-probs are float64 values around -900: their float32 copies are off by up to 3e-5 absolute, which
    # is a 3e-5 RELATIVE change of every softmax term
    assert abs(loss.item() - ref_loss) <= 1e-4 * abs(ref_loss)
    gr = lp.grad.double().cpu().numpy()
    assert np.abs(gr - g['grad']).max() <= 1e-4 * np.abs(g['grad']).max()
    assert np.all(gr[:, :-1] == 0.0)                      # only the last element of every sequence carries gradient


@pytest.mark.parametrize('M,B,T', [(1, 2, 1), (3, 5, 4), (40, 20, 2), (8, 70, 3)])
def test_episode_loss_kernel_matches_torch(M, B, T):
    from bruno_b200 import losses
    torch.manual_seed(M * 100 + B)
    lp = (torch.randn(M * B, T, device='cuda') * 20 - 900).requires_grad_(True)
    loss = losses.episode_softmax_loss(lp, M, B, T)
    (loss * 3.0).backward()
    lpd = lp.detach().double().requires_grad_(True)
    ref = -torch.log_softmax(lpd.reshape(M, B, T)[:, :, -1], dim=-1)[:, 0].mean()
    (ref * 3.0).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * max(abs(ref.item()), 1.0)
    assert (lp.grad.double() - lpd.grad).abs().max().item() <= 1e-5 * lpd.grad.abs().max().item() + 1e-9


def test_finetune_config_surface_and_gradient_flow():
    """The ft config keeps the reference's module attributes; the loss backward reaches every variable."""
    cfg = importlib.reload(importlib.import_module('bruno_b200.config_rnn.bn2_omniglot_tp_ft_1s_20w'))
    assert (cfg.batch_size, cfg.seq_len, cfg.meta_batch_size) == (20, 2, 8)
    assert cfg.learning_rate == 0.00003 and cfg.scale_student_grad == 1. and cfg.optimizer == 'rmsprop'
    x, y = next(cfg.train_data_iter.generate())
    assert x.shape == (160, 2, 28, 28, 1)
    x = torch.from_numpy(x[:40]).cuda()                   # two episodes
    cfg.meta_batch_size = 2
    with torch.no_grad():
        cfg.build_model(x[:1])
        for layer in cfg.model.layers():
            if hasattr(layer, 'Ks'):
                layer.Ks.normal_(0, 0.05)
                layer.Kt.normal_(0, 0.05)
    loss = cfg.loss(cfg.build_model(x)[0])
    grads = torch.autograd.grad(loss, cfg.model.parameters())
    assert np.isfinite(loss.item()) and 0.0 < loss.item()
    assert all(torch.isfinite(g).all() for g in grads)
    assert sum(float(g.abs().sum()) > 0 for g in grads) >= len(grads) - 4


def test_train_then_finetune_then_prior_samples(tmp_path):
    cwd = str(tmp_path)
    env = dict(os.environ, PYTHONPATH=ROOT)

    def run(mod, *args):
        p = subprocess.run([sys.executable, '-m', mod] + list(args), cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
        return p.stdout + p.stderr
    run('bruno_b200.config_rnn.train', '--config_name', 'bn2_omniglot_tp', '--nr_gpu', '1', '--max_iter', '3')
    out = run('bruno_b200.config_rnn.train_finetune', '--config_name', 'bn2_omniglot_tp_ft_1s_20w', '--max_iter', '4',
              '--print_every', '2')
    assert 'restoring parameters from' in out and 'train_loss=' in out and 'saving model' in out
    losses = [float(l.split('train_loss=')[1].split()[0]) for l in out.splitlines() if 'train_loss=' in l]
    assert all(np.isfinite(losses))
    ft = [d for d in os.listdir(os.path.join(cwd, 'metadata')) if d.startswith('bn2_omniglot_tp_ft_1s_20w-')]
    assert len(ft) == 1 and os.path.exists(os.path.join(cwd, 'metadata', ft[0], 'params.pt'))
    # prior samples of the base model (config_rnn/test_samples_prior.py)
    import shutil
    shutil.rmtree(os.path.join(cwd, 'metadata', ft[0]))   # find_model_metadata('bn2_omniglot_tp') must match one directory
    run('bruno_b200.config_rnn.test_samples_prior', '--config_name', 'bn2_omniglot_tp', '--n_row', '2', '--n_grids', '1')
    base = [d for d in os.listdir(os.path.join(cwd, 'metadata')) if d.startswith('bn2_omniglot_tp-')][0]
    s = np.load(os.path.join(cwd, 'metadata', base, 'samples', 'prior_sample_0.npy'))
    assert s.shape == (2, 2, 28, 28, 1) and np.isfinite(s).all()
    assert np.abs(s[0, 0] - s[0, 1]).max() > 0            # independent draws
