"""Generate tests/golden/ref_flow_*.npz by EXECUTING THE REFERENCE'S OWN, UNMODIFIED PYTHON.

Run in the build container only (needs /root/reference):   python tools/make_ref_flow_golden.py

What runs: /root/reference/{nn_extra_nvp,nn_extra_student,nn_extra_gauss,nn_extra_nvp_conditional}.py and the config
modules config_rnn/{bn2_omniglot_tp,bn2_omniglot_gp,en2_noconv_mnist_tp,an2_cifar_tp,bn2_omniglot_tp_ft_1s_20w}.py,
config_conditional/m1_shapenet.py -- imported as they are, over the eager float64 `tensorflow` stand-in of
tests/tf_shim/ (TF-1.7 cannot be installed here).  Full-size networks (64 filters, 8 (6) residual blocks, 512 (256)
units); only batch and sequence length are reduced.

Per config the fixture holds what the reference computes on seeded inputs / weights:
  * `init/*`    g and b of every weight-normalised layer after `build_model(x, init=True)` (nn_extra_nvp.py:394-398,
                422-426) -- the data-dependent initialisation -- from weights with RANDOMISED heads (so that s, t != 0 in
                the outputs / gradients below); `init0/*` the same from the reference's real iteration-0 state (heads = 0);
  * `log_probs`, `latent`, `prior`, `z_vec`   the four outputs of `build_model(x)` (bn2_omniglot_tp.py:59-168);
  * `loss`, `grad_*`   `config.loss(log_probs)` and d loss / d variable for every trainable variable through torch
                autograd over the reference code (what `tf.gradients(train_loss, all_params)` returns,
                config_rnn/train.py:91): [sum, L2 norm] of every gradient, the full gradient of every variable with at
                most 4096 elements, and a fixed strided sample of the larger ones;
  * `x_samples`, `sample_log_probs`, `noise_*`   the `sampling_mode=True` branch (bn2:106-158) with the uniforms /
                normals its samplers drew (logged by the stand-in so that the oracle and the GPU path consume the same).
Weights are NOT stored (14.7 M parameters): they are `oracle.init_params(spec, seed, perturb_process=True)` -- variable
names identical to the reference's (the generator asserts set equality) -- and the fixture keeps a checksum
(`weights_checksum`) so a drift of torch's CPU generator is detected instead of silently mis-comparing.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'tf_shim'))
import ref_loader                                  # noqa: E402
from oracle import bruno_oracle as O               # noqa: E402
from oracle import bruno_oracle_cond as OC         # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
SMALL = 256
NSAMPLE = 64

# config -> (B, T, weight seed, data kind)
CASES = {
    'bn2_omniglot_tp': (2, 3, 11, 'strokes'),
    'bn2_omniglot_gp': (2, 3, 12, 'strokes'),
    'en2_noconv_mnist_tp': (3, 4, 13, 'bytes'),
    'an2_cifar_tp': (2, 2, 14, 'bytes'),
}


def weights_checksum(P):
    s = np.zeros(3)
    for i, (k, v) in enumerate(P.items()):
        v = v.double()
        s += np.array([v.sum().item(), (v * v).sum().item(), (v.flatten()[::7] * (1 + i % 5)).sum().item()])
    return s


def grad_summary(name, g):
    g = g.detach().double().reshape(-1)
    out = {'gsum/' + name: np.array([g.sum().item(), g.norm().item()])}
    if g.numel() <= SMALL:
        out['gfull/' + name] = g.numpy().copy()
    else:
        out['gsamp/' + name] = g[sample_index(g.numel())].numpy().copy()
    return out


def sample_index(n):
    return torch.linspace(0, n - 1, NSAMPLE).round().long()


def wn_names(P):
    return [k for k in P if k.endswith('/g') or k.endswith('/b')]


def run_unconditional(name):
    B, T, seed, kind = CASES[name]
    spec = O.spec_for(name)
    P0 = O.init_params(spec, seed=seed, dtype=torch.float64, perturb_process=True)
    x = O.synthetic_batch(spec, B, T, seed=seed + 1, kind=kind)
    fix = {'config': name, 'B': B, 'T': T, 'seed': seed, 'kind': kind, 'weights_checksum': weights_checksum(P0)}

    # (1) data-dependent init + forward, one "session"
    with ref_loader.reference_modules() as tf:
        tf.preload_variables({k: v.numpy() for k, v in P0.items()})
        cfg = ref_loader.import_config('config_rnn', name, seq_len=T)
        with torch.no_grad(), tf.variable_scope('model'):
            cfg.build_model(tf._f(x), init=True)
            assert set(tf.variables_by_name()) == set(P0), 'variable names differ from SURVEY 9.8 / oracle.init_params'
            assert all(tuple(v.shape) == tuple(P0[k].shape) for k, v in tf.variables_by_name().items())
            P1 = {k: v.detach().clone() for k, v in tf.variables_by_name().items()}
    for k in wn_names(P0):
        fix['init/' + k] = P1[k].numpy()
    for k in P0:
        if k not in wn_names(P0):
            assert torch.equal(P1[k], P0[k]), k            # init touches only g and b

    # (1b) the same init pass from the reference's REAL iteration-0 state: zero-initialised heads (nvp:129-138), i.e. the
    #      flow is the identity and every coupling's s/t network sees the exact logit-transformed input
    Pz = O.init_params(spec, seed=seed, dtype=torch.float64, randomize_heads=False, perturb_process=True)
    with ref_loader.reference_modules() as tf:
        tf.preload_variables({k: v.numpy() for k, v in Pz.items()})
        cfg = ref_loader.import_config('config_rnn', name, seq_len=T)
        with torch.no_grad(), tf.variable_scope('model'):
            cfg.build_model(tf._f(x), init=True)
            for k, v in tf.variables_by_name().items():
                if k.endswith('/g') or k.endswith('/b'):
                    fix['init0/' + k] = v.detach().numpy().copy()

    # (2) forward + loss + gradients with the initialised weights (fresh graph: the student layer derives var/nu/corr
    #     from its variables at construction)
    with ref_loader.reference_modules() as tf:
        tf.preload_variables({k: v.numpy() for k, v in P1.items()})
        cfg = ref_loader.import_config('config_rnn', name, seq_len=T)
        with tf.variable_scope('model'):
            lp, lat, pri, z = cfg.build_model(tf._f(x))
        loss = cfg.loss(lp)
        vars_ = tf.variables_by_name()
        names = [k for k in vars_ if k in P0]
        grads = torch.autograd.grad(loss, [vars_[k] for k in names], allow_unused=True)
        fix.update(log_probs=lp.detach().numpy(), latent=lat.detach().numpy(), prior=pri.detach().numpy(),
                   z_vec=z.detach().numpy(), loss=float(loss))
        for k, g in zip(names, grads):
            fix.update(grad_summary(k, torch.zeros_like(vars_[k]) if g is None else g))

    # (3) sampling branch: sample_batch_size = 1 sequence (config attribute), n_samples = 4
    with ref_loader.reference_modules() as tf:
        tf.preload_variables({k: v.numpy() for k, v in P1.items()})
        cfg = ref_loader.import_config('config_rnn', name, seq_len=T)
        with torch.no_grad(), tf.variable_scope('model'):
            xs, slp = cfg.build_model(tf._f(x[:cfg.sample_batch_size]), sampling_mode=True)
        fix.update(x_samples=xs.numpy(), sample_log_probs=slp.numpy(), n_samples=cfg.n_samples)
        for i, r in enumerate(tf.random_log()):
            fix['noise/%02d' % i] = r.numpy()
    np.savez_compressed(os.path.join(OUT, 'ref_flow_%s.npz' % name), **fix)
    print(name, 'loss %.6f' % fix['loss'], 'x_samples', fix['x_samples'].shape, '%d keys' % len(fix))


def run_conditional():
    name, B, T, seed, n_context = 'm1_shapenet', 2, 3, 15, 1
    spec = OC.CondSpec()
    P0 = OC.init_params(spec, seed=seed, dtype=torch.float64, perturb_process=True)
    x, y = OC.synthetic_batch(spec, B, T, seed=seed + 1)
    fix = {'config': name, 'B': B, 'T': T, 'seed': seed, 'n_context': n_context, 'weights_checksum': weights_checksum(P0)}
    with ref_loader.reference_modules() as tf:
        tf.preload_variables({k: v.numpy() for k, v in P0.items()})
        cfg = ref_loader.import_config('config_conditional', name, seq_len=T, n_context=n_context)
        with torch.no_grad(), tf.variable_scope('model'):
            cfg.build_model(tf._f(x), tf._f(y), init=True)
            got = set(tf.variables_by_name())
            assert got == set(P0), (sorted(got - set(P0))[:5], sorted(set(P0) - got)[:5])
            P1 = {k: v.detach().clone() for k, v in tf.variables_by_name().items()}
    for k in wn_names(P0):
        fix['init/' + k] = P1[k].numpy()
    with ref_loader.reference_modules() as tf:
        tf.preload_variables({k: v.numpy() for k, v in P1.items()})
        cfg = ref_loader.import_config('config_conditional', name, seq_len=T, n_context=n_context)
        with tf.variable_scope('model'):
            lp = cfg.build_model(tf._f(x), tf._f(y))[0]
        loss = cfg.loss(lp)
        vars_ = tf.variables_by_name()
        names = [k for k in vars_ if k in P0]
        trainable = {v.var_name for v in tf.trainable_variables()}
        grads = torch.autograd.grad(loss, [vars_[k] for k in names], allow_unused=True)
        fix.update(log_probs=lp.detach().numpy(), loss=float(loss),
                   non_trainable=np.array(sorted(set(names) - trainable)))
        for k, g in zip(names, grads):
            fix.update(grad_summary(k, torch.zeros_like(vars_[k]) if g is None else g))
    with ref_loader.reference_modules() as tf:
        tf.preload_variables({k: v.numpy() for k, v in P1.items()})
        cfg = ref_loader.import_config('config_conditional', name, seq_len=T, n_context=n_context)
        with torch.no_grad(), tf.variable_scope('model'):
            xs = cfg.build_model(tf._f(x), tf._f(y), sampling_mode=True)
        fix.update(x_samples=xs.numpy())
        for i, r in enumerate(tf.random_log()):
            fix['noise/%02d' % i] = r.numpy()
    np.savez_compressed(os.path.join(OUT, 'ref_flow_%s.npz' % name), **fix)
    print(name, 'loss %.6f' % fix['loss'], 'x_samples', fix['x_samples'].shape, '%d keys' % len(fix))


def run_finetune_loss():
    """The episodic fine-tuning loss of config_rnn/bn2_omniglot_tp_ft_1s_20w.py:222-226 on seeded log-probs."""
    rs = np.random.RandomState(5)
    with ref_loader.reference_modules() as tf:
        cfg = ref_loader.import_config('config_rnn', 'bn2_omniglot_tp_ft_1s_20w', seq_len=2)
        M, Bf, T = cfg.meta_batch_size, cfg.batch_size, cfg.seq_len
        lp = torch.from_numpy(rs.normal(-900.0, 15.0, size=(M * Bf, T))).requires_grad_(True)
        loss = cfg.loss(tf._t(lp))
        g, = torch.autograd.grad(loss, lp)
    np.savez_compressed(os.path.join(OUT, 'ref_finetune_loss.npz'), log_probs=lp.detach().numpy(), loss=float(loss),
                        grad=g.numpy(), meta_batch_size=M, batch_size=Bf, seq_len=T)
    print('finetune loss %.6f' % float(loss), (M, Bf, T))


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count())
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or list(CASES) + ['m1_shapenet', 'finetune']
    for nm in which:
        if nm == 'm1_shapenet':
            run_conditional()
        elif nm == 'finetune':
            run_finetune_loss()
        else:
            run_unconditional(nm)
