"""GPU parity: bruno_process_* (through the C ABI) vs the CPU oracle and the reference's golden vectors.
Tolerance (north star): per-example log-likelihood within 1e-4 relative in fp32."""
import os

import numpy as np
import pytest
import torch

from oracle import bruno_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def _layer(kind, D, seed=0, perturb=True, mu=0.):
    from bruno_b200.nn_extra_student import StudentRecurrentLayer
    from bruno_b200.nn_extra_gauss import GaussianRecurrentLayer
    spec = O.Spec('t', (1, 1, D), 1, [], [], 1, kind)
    P = O.init_params(spec, seed=seed, dtype=torch.float64, perturb_process=perturb)
    if kind == 'tp':
        layer = StudentRecurrentLayer((D,), mu_init=mu)
        layer.nu_vbl.data.copy_(P['model/prior_nu'].float())
        layer.var_vbl.data.copy_(P['model/prior_var'].float())
        layer.corr_vbl.data.copy_(P['model/prior_corr'].float())
    else:
        layer = GaussianRecurrentLayer((D,), mu_init=mu)
        layer.var_vbl.data.copy_(P['model/gaussian/prior_var'].float())
        layer.corr_vbl.data.copy_(P['model/gaussian/prior_corr'].float())
    # the oracle sees exactly the float32-rounded variables the GPU sees
    P = {k: v.float().double() for k, v in P.items()}
    return layer, P


def _oracle_proc(kind, P, mu=0.):
    D = P['model/prior_var' if kind == 'tp' else 'model/gaussian/prior_var'].shape[1]
    m = torch.full((1, D), float(mu), dtype=torch.float64)
    return O.StudentOracle(P, mu=m) if kind == 'tp' else O.GaussOracle(P, mu=m)


def _close(a, b, rtol=RTOL, atol=1e-5):
    a = a.detach().double().cpu()
    assert torch.allclose(a, b, rtol=rtol, atol=atol), (a - b).abs().max().item()


def test_reference_golden_vectors(golden_dir):
    """The three checks of the reference's tests_gp_tp/test_layers.py (:127/:141, :161, :176), same tolerance."""
    from bruno_b200.nn_extra_student import StudentRecurrentLayer
    from bruno_b200.nn_extra_gauss import GaussianRecurrentLayer
    g = np.load(os.path.join(golden_dir, 'tp_gp_reference.npz'))
    x = torch.from_numpy(g['x']).float().cuda()            # [1,100,1]
    tp = StudentRecurrentLayer((1,), nu_init=float(g['nu'][0]), var_init=float(g['var'][0]), corr_init=float(g['corr'][0]))
    gp = GaussianRecurrentLayer((1,), var_init=float(g['var'][0]), corr_init=float(g['corr'][0]))
    lp, _ = tp.sequence_log_likelihoods(x)
    lp = lp.detach().cpu().numpy()[0]
    assert np.allclose(lp, g['tp_recursive'], atol=1e-3, rtol=1e-3)
    assert np.allclose(lp[1:], g['tp_joint'][1:], atol=1e-3, rtol=1e-3)
    lg, _ = gp.sequence_log_likelihoods(x)
    assert np.allclose(lg.detach().cpu().numpy()[0][1:], g['gp_joint'][1:], atol=1e-3, rtol=1e-3)


@pytest.mark.parametrize('kind', ['tp', 'gp'])
@pytest.mark.parametrize('B,T,D', [(5, 20, 784), (32, 20, 784), (3, 2, 3072), (2, 7, 6), (1, 1, 4), (20, 2, 784), (4, 100, 1024)])
def test_forward_matches_oracle(kind, B, T, D):
    layer, P = _layer(kind, D, seed=B + T)
    g = torch.Generator().manual_seed(D + T)
    z = torch.randn(B, T, D, generator=g, dtype=torch.float64) * 1.5 + 0.3 * torch.randn(B, 1, D, generator=g, dtype=torch.float64)
    z = z.float()
    lp_o, pr_o = O.process_log_probs(_oracle_proc(kind, P), z.double())
    lp, pr = layer.sequence_log_likelihoods(z.cuda())
    _close(lp, lp_o)
    _close(pr, pr_o)
    lp2, none = layer.sequence_log_likelihoods(z.cuda(), want_prior=False)
    assert none is None
    assert torch.equal(lp2, lp)


@pytest.mark.parametrize('kind', ['tp', 'gp'])
def test_nonzero_mean_and_mask_dim(kind):
    B, T, D = 3, 5, 64
    layer, P = _layer(kind, D, seed=11, mu=0.7)
    z = torch.randn(B, T, D, generator=torch.Generator().manual_seed(5)).float()
    mask = (torch.arange(D) % 3 != 0).double().view(1, D)
    lp_o, pr_o = O.process_log_probs(_oracle_proc(kind, P, mu=0.7), z.double(), mask_dim=mask)
    lp, pr = layer.sequence_log_likelihoods(z.cuda(), mask_dim=mask.float().cuda())
    _close(lp, lp_o)
    _close(pr, pr_o)


@pytest.mark.parametrize('kind', ['tp', 'gp'])
def test_stepwise_protocol_matches_fused(kind):
    """reset / get_log_likelihood / get_log_likelihood_under_prior / update_distribution (the reference's API)."""
    B, T, D = 4, 6, 784
    layer, P = _layer(kind, D, seed=2)
    z = torch.randn(B, T, D, generator=torch.Generator().manual_seed(9)).float().cuda()
    lp, pr = layer.sequence_log_likelihoods(z)
    layer.reset()
    for t in range(T):
        a = layer.get_log_likelihood(z[:, t, :])
        b = layer.get_log_likelihood_under_prior(z[:, t, :])
        assert torch.allclose(a, lp[:, t].detach(), rtol=1e-5, atol=1e-3)
        assert torch.allclose(b, pr[:, t].detach(), rtol=1e-5, atol=1e-3)
        layer.update_distribution(z[:, t, :])
    proc = _oracle_proc(kind, P)
    for t in range(T):
        proc.update_distribution(z[:, t, :].double().cpu())
    cur = layer.current_distribution
    _close(cur.mu, proc.cur[0], rtol=1e-4, atol=1e-5)
    _close(cur.var, proc.cur[1], rtol=1e-4, atol=1e-5)
    if kind == 'tp':
        _close(cur.nu, proc.cur[2], rtol=1e-5)


@pytest.mark.parametrize('kind', ['tp', 'gp'])
@pytest.mark.parametrize('B,T,D,with_prior', [(5, 20, 784, False), (3, 4, 3072, True), (2, 5, 6, True), (32, 20, 784, False)])
def test_backward_matches_oracle_autograd(kind, B, T, D, with_prior):
    layer, P = _layer(kind, D, seed=3)
    g = torch.Generator().manual_seed(21)
    z = (torch.randn(B, T, D, generator=g) * 1.2).float()
    wl = torch.randn(B, T, generator=g).double()
    wp = torch.randn(B, T, generator=g).double()
    # oracle
    keys = ['model/prior_var', 'model/prior_nu', 'model/prior_corr'] if kind == 'tp' else ['model/gaussian/prior_var', 'model/gaussian/prior_corr']
    leaves = {k: P[k].clone().requires_grad_(True) for k in keys}
    zo = z.double().requires_grad_(True)
    lp_o, pr_o = O.process_log_probs(_oracle_proc(kind, leaves), zo)
    obj = (lp_o * wl).sum() + ((pr_o * wp).sum() if with_prior else 0.)
    grads = torch.autograd.grad(obj, [zo] + [leaves[k] for k in keys])
    # gpu
    zg = z.cuda().requires_grad_(True)
    lp, pr = layer.sequence_log_likelihoods(zg, want_prior=with_prior)
    objg = (lp * wl.float().cuda()).sum() + ((pr * wp.float().cuda()).sum() if with_prior else 0.)
    params = [layer.var_vbl, layer.nu_vbl, layer.corr_vbl] if kind == 'tp' else [layer.var_vbl, layer.corr_vbl]
    gg = torch.autograd.grad(objg, [zg] + params)
    for a, b in zip(gg, grads):
        a = a.double().cpu()
        scale = b.abs().max().item() + 1e-12
        assert (a - b).abs().max().item() <= 2e-3 * scale, ((a - b).abs().max().item(), scale)


def test_sampler_matches_oracle():
    B, T, D, n = 1, 5, 784, 4
    for kind in ('tp', 'gp'):
        layer, P = _layer(kind, D, seed=4)
        g = torch.Generator().manual_seed(1)
        z = torch.randn(B, T, D, generator=g).float()
        mu, var, nu = layer.posterior_sequence(z.cuda())
        proc = _oracle_proc(kind, P)
        for t in range(T + 1):
            noise = torch.rand(2, n, B, D, generator=g) if kind == 'tp' else torch.randn(n, B, D, generator=g)
            ref = proc.sample_from_uniforms(noise.double()) if kind == 'tp' else proc.sample_from_normals(noise.double())
            out = layer.sample_from(mu[:, t], var[:, t], None if nu is None else nu[t], noise.cuda())
            _close(out, ref, rtol=5e-4, atol=5e-4)   # fp32 sin/cos/sqrt chain of the polar method
            if t < T:
                proc.update_distribution(z[:, t, :].double())


@pytest.mark.parametrize('kind', ['tp', 'gp'])
def test_exchangeability_at_scale(kind):
    """Size-independent property at an HBM-resident size: the joint log-likelihood sum_t lp[b,t] of an exchangeable
    process does not depend on the order of the sequence."""
    B, T, D = 2048, 20, 784
    layer, _ = _layer(kind, D, seed=8)
    z = torch.randn(B, T, D, device='cuda')
    perm = torch.randperm(T, device='cuda')
    lp, _ = layer.sequence_log_likelihoods(z)
    lq, _ = layer.sequence_log_likelihoods(z[:, perm, :].contiguous())
    a, b = lp.sum(1).double(), lq.sum(1).double()
    assert torch.allclose(a, b, rtol=2e-5, atol=1e-2), (a - b).abs().max().item()


@pytest.mark.parametrize('kind', ['tp', 'gp'])
def test_backward_is_finite_for_huge_latents(kind):
    """Untrained flows emit |z| ~ 1e4: the reverse pass must not lose the prefix sums (q = 1 + e S2 + f S1^2 > 0)."""
    B, T, D = 2, 3, 64
    layer, P = _layer(kind, D, seed=5)
    g = torch.Generator().manual_seed(2)
    z = (torch.randn(B, T, D, generator=g) * 3e3 + 1.2e4).float()
    keys = ['model/prior_var', 'model/prior_nu', 'model/prior_corr'] if kind == 'tp' else ['model/gaussian/prior_var', 'model/gaussian/prior_corr']
    leaves = {k: P[k].clone().requires_grad_(True) for k in keys}
    zo = z.double().requires_grad_(True)
    lp_o, _ = O.process_log_probs(_oracle_proc(kind, leaves), zo)
    grads = torch.autograd.grad(lp_o.sum(), [zo] + [leaves[k] for k in keys])
    zg = z.cuda().requires_grad_(True)
    lp, _ = layer.sequence_log_likelihoods(zg)
    params = [layer.var_vbl, layer.nu_vbl, layer.corr_vbl] if kind == 'tp' else [layer.var_vbl, layer.corr_vbl]
    gg = torch.autograd.grad(lp.sum(), [zg] + params)
    _close(lp, lp_o, rtol=1e-4, atol=1e-2)
    for a, b in zip(gg, grads):
        a = a.double().cpu()
        assert torch.isfinite(a).all()
        assert (a - b).abs().max().item() <= 5e-3 * (b.abs().max().item() + 1e-12)


@pytest.mark.parametrize('kind', ['tp', 'gp'])
@pytest.mark.parametrize('B,T,D,mu', [(7, 20, 784, 0.), (5, 7, 3600, 0.), (9, 3, 40, 0.4), (1030, 20, 784, 0.)])
def test_forward_scan_variants_agree(kind, B, T, D, mu):
    """The persistent bulk-copy ring kernel (default, shapes 10-15) against the register-staged scan (9) and the
    oracle: ragged row groups, a ragged last segment (3600 = 4 x 896 + 16), non-zero mu0, more groups than resident
    blocks, and the running-state entry (s1/s2 read and written back)."""
    from bruno_b200 import _lib
    from bruno_b200.process import build_table
    layer, P = _layer(kind, D, seed=3, mu=mu)
    g = torch.Generator().manual_seed(B + D)
    z = (torch.randn(B, T, D, generator=g) * 1.3 + 0.5 * torch.randn(B, 1, D, generator=g)).float().cuda()
    out = {}
    try:
        for v in (9, 0, 10, 11, 12, 13, 14, 15, 16, 17, 18):
            _lib.query('bruno_process_set_fwd_variant', v)
            lp, pr = layer.sequence_log_likelihoods(z)
            out[v] = (lp.detach().clone(), pr.detach().clone())
        for v in (0, 10, 11, 12, 13, 14, 15, 16, 17, 18):
            assert torch.allclose(out[v][0], out[9][0], rtol=3e-5, atol=1e-3), (v, (out[v][0] - out[9][0]).abs().max().item())
            assert torch.allclose(out[v][1], out[9][1], rtol=3e-5, atol=1e-3), v
        if B <= 16:
            lp_o, pr_o = O.process_log_probs(_oracle_proc(kind, P, mu=mu), z.double().cpu())
            _close(out[0][0], lp_o)
            _close(out[0][1], pr_o)
        # running state: the first T1 steps, then the rest from the stored prefix sums
        T1 = T // 2
        if T1 > 0:
            nu = layer.nu_vbl.detach() if kind == 'tp' else None
            for v in (9, 0):
                _lib.query('bruno_process_set_fwd_variant', v)
                s1 = torch.zeros(B, D, device='cuda')
                s2 = torch.zeros(B, D, device='cuda')
                lps = []
                for n0, zz in ((0, z[:, :T1].contiguous()), (T1, z[:, T1:].contiguous())):
                    Tn = zz.shape[1]
                    tab = build_table(layer.kind, layer.var_vbl.detach(), nu, layer.corr_vbl.detach(), layer.mu, None,
                                      layer.min_nu, n0, Tn, False)
                    lp = torch.empty(B, Tn, device='cuda')
                    scratch = torch.empty(max(_lib.query('bruno_process_scratch_floats', B, Tn, D), 1), device='cuda')
                    _lib.call('bruno_process_fwd', layer.kind, zz, layer.mu, tab, s1, s2, lp, None, scratch, B, Tn, D)
                    lps.append(lp)
                lp_cat = torch.cat(lps, 1)
                assert torch.allclose(lp_cat, out[9][0], rtol=3e-5, atol=2e-3), (v, (lp_cat - out[9][0]).abs().max().item())
    finally:
        _lib.query('bruno_process_set_fwd_variant', 0)


@pytest.mark.parametrize('kind', ['tp', 'gp'])
@pytest.mark.parametrize('B,T,D,mu,with_prior', [(7, 20, 784, 0., False), (5, 7, 3600, 0., True), (9, 3, 40, 0.4, True),
                                                 (1190, 6, 784, 0., False), (3, 1, 784, 0., True)])
def test_backward_scan_variants_agree(kind, B, T, D, mu, with_prior):
    """The bulk-copy ring backward (default, shapes 10-12) against the register-staged kernel (9): ragged groups, a
    ragged last segment, non-zero mu0, more groups than persistent blocks, T = 1, with and without the prior term."""
    from bruno_b200 import _lib
    layer, P = _layer(kind, D, seed=4, mu=mu)
    g = torch.Generator().manual_seed(B * 7 + D)
    z = (torch.randn(B, T, D, generator=g) * 1.2 + 0.4 * torch.randn(B, 1, D, generator=g)).float().cuda()
    wl = torch.randn(B, T, generator=g).cuda()
    wp = torch.randn(B, T, generator=g).cuda()
    params = [layer.var_vbl, layer.nu_vbl, layer.corr_vbl] if kind == 'tp' else [layer.var_vbl, layer.corr_vbl]
    out = {}
    try:
        for v in (9, 0, 10, 11, 12, 13, 14, 15):
            _lib.query('bruno_process_set_bwd_variant', v)
            zg = z.clone().requires_grad_(True)
            lp, pr = layer.sequence_log_likelihoods(zg, want_prior=with_prior)
            obj = (lp * wl).sum() + ((pr * wp).sum() if with_prior else 0.)
            out[v] = [t.detach().clone() for t in torch.autograd.grad(obj, [zg] + params)]
        for v in (0, 10, 11, 12, 13, 14, 15):
            for a, b in zip(out[v], out[9]):
                scale = b.abs().max().item() + 1e-12
                assert (a - b).abs().max().item() <= 1e-4 * scale, (v, (a - b).abs().max().item(), scale)
    finally:
        _lib.query('bruno_process_set_bwd_variant', 0)
