"""CPU: pins the oracle's Real NVP / build_model restatement to the REFERENCE'S OWN CODE.

tests/golden/ref_flow_*.npz were written by tools/make_ref_flow_golden.py, which imports the unmodified reference
modules (/root/reference/nn_extra_nvp.py, nn_extra_student.py, nn_extra_gauss.py, nn_extra_nvp_conditional.py,
config_rnn/*.py, config_conditional/m1_shapenet.py) over the eager float64 `tensorflow` stand-in in tests/tf_shim/ and
records data-dependent init, the four build_model outputs, loss + gradients and the sampling branch of FULL-SIZE
networks.  Here the oracle must reproduce every number to float64 rounding (1e-9 relative) -- masks, layer order,
squeeze / factor-out, weight norm, log-det, inverse, init, variable names are then pinned to the reference, not to our
reading of it.  The last test re-executes the reference live (build container only; skipped on the GPU box).
"""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import bruno_oracle as O
from oracle import bruno_oracle_cond as OC

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), 'tf_shim'))
import ref_loader                       # noqa: E402

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tools'))
from make_ref_flow_golden import weights_checksum, sample_index      # noqa: E402

RTOL = 1e-9
UNCOND = ['bn2_omniglot_tp', 'bn2_omniglot_gp', 'en2_noconv_mnist_tp', 'an2_cifar_tp']


def close(a, b, scale=None, rtol=RTOL):
    """max |a - b| <= rtol * scale on the finite entries; non-finite entries (the reference's own fp overflow in
    LogitLayer.backward, nn_extra_nvp.py:36, on untrained an2 samples) must be non-finite in the same places."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    fin = np.isfinite(b)
    if not np.array_equal(fin, np.isfinite(a)):
        return False, np.inf
    if not fin.any():
        return True, 0.0
    s = max(np.abs(b[fin]).max(), 1e-300) if scale is None else scale
    e = np.abs(a[fin] - b[fin]).max()
    return e <= rtol * s, e / s


def load(golden_dir, name):
    g = np.load(os.path.join(golden_dir, 'ref_flow_%s.npz' % name))
    return g, int(g['B']), int(g['T']), int(g['seed'])


def initialised(g, P):
    for k in P:
        if k.endswith('/g') or k.endswith('/b'):
            P[k] = torch.from_numpy(g['init/' + k]).clone()
    return P


def check_grads(g, grads):
    """Every gradient vs the reference's: sum, L2 norm and the stored entries.  Tolerances are relative to the
    variable's own gradient norm, floored at 1e-6 of the largest gradient norm of the model (gradients that are
    analytically zero -- e.g. d loss / d V along V under weight norm -- are pure rounding noise on both sides)."""
    G = max(float(g[k][1]) for k in g.files if k.startswith('gsum/'))
    worst = 0.0
    for k, gr in grads.items():
        gr = gr.reshape(-1).double()
        ref = g['gsum/' + k]
        scale = max(ref[1], 1e-6 * G)
        assert abs(gr.sum().item() - ref[0]) <= 1e-7 * scale * np.sqrt(gr.numel()), k
        assert abs(gr.norm().item() - ref[1]) <= 1e-8 * scale, k
        if 'gfull/' + k in g.files:
            ok, e = close(gr.numpy(), g['gfull/' + k], scale=scale, rtol=1e-7)
        else:
            ok, e = close(gr[sample_index(gr.numel())].numpy(), g['gsamp/' + k], scale=scale, rtol=1e-7)
        assert ok, (k, e)
        worst = max(worst, e)
    return worst


@pytest.mark.parametrize('name', UNCOND)
def test_oracle_equals_reference_code(golden_dir, name):
    g, B, T, seed = load(golden_dir, name)
    spec = O.spec_for(name)
    P = O.init_params(spec, seed=seed, dtype=torch.float64, perturb_process=True)
    assert np.allclose(weights_checksum(P), g['weights_checksum'], rtol=1e-12), 'torch CPU generator drifted: regenerate'
    x = O.synthetic_batch(spec, B, T, seed=seed + 1, kind=str(g['kind']))
    # data-dependent init (nn_extra_nvp.py:394-398, 422-426)
    O.data_dependent_init(spec, P, x)
    for k in P:
        if k.endswith('/g') or k.endswith('/b'):
            ok, e = close(P[k].numpy(), g['init/' + k])
            assert ok, (k, e)
    P = initialised(g, P)
    # build_model outputs (bn2_omniglot_tp.py:59-168)
    with torch.no_grad():
        lp, lat, pri, z = O.build_model(spec, P, x)
    for got, key in ((lp, 'log_probs'), (lat, 'latent'), (pri, 'prior'), (z, 'z_vec')):
        ok, e = close(got.numpy(), g[key])
        assert ok, (key, e)
    # loss + gradients (train.py:88-91)
    loss, grads = O.loss_and_grads(spec, P, x)
    assert abs(loss.item() - float(g['loss'])) <= RTOL * abs(float(g['loss']))
    check_grads(g, grads)
    # sampling branch (bn2:106-158) with the reference's own random draws
    noise = [torch.from_numpy(g[k]) for k in sorted(k for k in g.files if k.startswith('noise/'))]
    assert len(noise) == T + 1
    with torch.no_grad():
        xs, slp = O.build_model_sampling(spec, P, x[:1], noise)
    ok, e = close(xs.numpy(), g['x_samples'], scale=256.0, rtol=1e-8)
    assert ok, e
    ok, e = close(slp.numpy(), g['sample_log_probs'])
    assert ok, e


def test_conditional_oracle_equals_reference_code(golden_dir):
    g, B, T, seed = load(golden_dir, 'm1_shapenet')
    nc = int(g['n_context'])
    spec = OC.CondSpec()
    P = OC.init_params(spec, seed=seed, dtype=torch.float64, perturb_process=True)
    assert np.allclose(weights_checksum(P), g['weights_checksum'], rtol=1e-12)
    x, y = OC.synthetic_batch(spec, B, T, seed=seed + 1)
    OC.data_dependent_init(spec, P, x, y, n_context=nc)
    for k in P:
        if k.endswith('/g') or k.endswith('/b'):
            ok, e = close(P[k].numpy(), g['init/' + k])
            assert ok, (k, e)
    P = initialised(g, P)
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    lp = OC.build_model(spec, leaves, x, y, n_context=nc)[0]
    ok, e = close(lp.detach().numpy(), g['log_probs'])
    assert ok, e
    loss = O.loss(lp)
    assert abs(loss.item() - float(g['loss'])) <= RTOL * abs(float(g['loss']))
    grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
    grads = {k: (torch.zeros_like(v) if gr is None else gr) for (k, v), gr in zip(leaves.items(), grads)}
    check_grads(g, grads)
    # `use_bias=False` only makes b non-trainable; the variable exists and is added (cond:444-445)
    assert all(k.endswith('/b') for k in g['non_trainable'])
    noise = [torch.from_numpy(g[k]) for k in sorted(k for k in g.files if k.startswith('noise/'))]
    with torch.no_grad():
        xs = OC.build_model_sampling(spec, P, x, y, noise, n_context=nc)
    ok, e = close(xs.reshape(g['x_samples'].shape).numpy(), g['x_samples'], scale=1.0, rtol=1e-8)
    assert ok, e


def test_finetune_loss_equals_reference_code(golden_dir):
    """config_rnn/bn2_omniglot_tp_ft_1s_20w.py:222-226."""
    g = np.load(os.path.join(golden_dir, 'ref_finetune_loss.npz'))
    lp = torch.from_numpy(g['log_probs']).requires_grad_(True)
    loss = O.finetune_loss(lp, int(g['meta_batch_size']), int(g['batch_size']), int(g['seq_len']))
    gr, = torch.autograd.grad(loss, lp)
    assert abs(loss.item() - float(g['loss'])) <= 1e-12 * abs(float(g['loss']))
    assert np.allclose(gr.numpy(), g['grad'], rtol=1e-10, atol=1e-14)


@pytest.mark.skipif(not ref_loader.available(), reason='reference tree only exists in the build container')
def test_live_reference_execution_matches_oracle():
    """Runs the unmodified reference build_model live on a seed no fixture uses (few-shot shape B=3, T=2)."""
    name, B, T, seed = 'bn2_omniglot_tp', 3, 2, 101
    spec = O.spec_for(name)
    P = O.init_params(spec, seed=seed, dtype=torch.float64, perturb_process=True)
    x = O.synthetic_batch(spec, B, T, seed=seed + 1)
    with ref_loader.reference_modules() as tf:
        tf.preload_variables({k: v.numpy() for k, v in P.items()})
        cfg = ref_loader.import_config('config_rnn', name, seq_len=T)
        with torch.no_grad(), tf.variable_scope('model'):
            ref = [torch.as_tensor(t).clone() for t in cfg.build_model(tf._f(x))]
        assert set(tf.variables_by_name()) == set(P)
    with torch.no_grad():
        got = O.build_model(spec, P, x)
    for a, b in zip(got, ref):
        assert close(a.numpy(), b.numpy())[0]
