"""GPU parity of the whole hot path (flow + process, forward and hand-written backward, inverse / sampling) against
the CPU oracle on identical synthetic inputs and weights.

Tolerances.  Dense (en2) path is fp32 end to end: per-example log-likelihood within 1e-4 relative (north star).
Conv path: bf16 tensor-core operands with fp32 accumulation and bf16 hidden activations -> stated looser bound:
per-example log-likelihood within 6e-3 relative (measured 0.6e-3 .. 3.9e-3 on these untrained, randomly initialised
networks whose |log p| is ~1.6e4 nats per frame), gradients within 1e-1 of each tensor's max (measured <= 5.3e-2)."""
import importlib

import numpy as np
import pytest
import torch

from oracle import bruno_oracle as O

pytestmark = pytest.mark.gpu

# against the oracle run WITH the declared bf16 storage points emulated (oracle.emulate_bf16).  One conv layer then
# agrees bit for bit except ~0.01-0.03 % one-ulp flips (fp32 accumulation order), but every flipped value perturbs 576
# downstream sums and re-flips a few of them: the mismatch fraction grows ~2.5x per layer and saturates near the bf16
# noise floor after ~16 layers (tools/dbg_coupling.py, DESIGN.md "Precision").  So the whole-model bound vs the
# emulating oracle is only ~2-4x tighter than vs the exact oracle; the bitwise statement is made per layer below.
TOL_LL_EMU = 1.5e-3
TOL_GRAD_EMU = 3e-2
TOL_LL = {'en2_noconv_mnist_tp': 1e-4, 'bn2_omniglot_tp': 6e-3, 'bn2_omniglot_gp': 6e-3, 'an2_cifar_tp': 6e-3}
AN2_TOL_GRAD_EXACT, AN2_TOL_GRAD_EMU = 5e-2, 3e-2     # measured on a B200: 1.6e-2 / 9.1e-3 of each tensor's max (smooth input)
TOL_GRAD = {'en2_noconv_mnist_tp': 2e-3, 'bn2_omniglot_tp': 1e-1, 'bn2_omniglot_gp': 1e-1, 'an2_cifar_tp': 1.0}


def fresh_config(name):
    mod = importlib.import_module('bruno_b200.config_rnn.' + name)
    return importlib.reload(mod)


def setup(name, B, T, seed, kind=None):
    spec = O.spec_for(name)
    P = O.init_params(spec, seed=seed, dtype=torch.float64, perturb_process=True)
    x = O.synthetic_batch(spec, B, T, seed=seed + 1, kind=kind or ('strokes' if 'cifar' not in name else 'bytes'))
    x = x.float().double()
    O.data_dependent_init(spec, P, x)
    P = O.cast_params(O.cast_params(P, torch.float32), torch.float64)   # both sides see the same fp32 numbers
    cfg = fresh_config(name)
    xg = x.float().cuda()
    cfg.build_model(xg[:1])              # creates the variables
    cfg.model.load_tf_variables(P)
    return spec, P, x, cfg, xg


def rel_err(a, b):
    a = a.detach().double().cpu()
    return ((a - b).abs() / b.abs().clamp_min(1.0)).max().item()


@pytest.mark.parametrize('name,B,T', [('en2_noconv_mnist_tp', 4, 5), ('bn2_omniglot_tp', 2, 3), ('bn2_omniglot_gp', 2, 2),
                                      ('an2_cifar_tp', 1, 2)])
def test_forward_and_backward_match_oracle(name, B, T):
    spec, P, x, cfg, xg = setup(name, B, T, seed=5)
    with torch.no_grad():
        lp_o, lat_o, pri_o, z_o = O.build_model(spec, P, x)
    lp, lat, pri, z = cfg.build_model(xg)
    e = {k: rel_err(a, b) for k, a, b in (('log_probs', lp, lp_o), ('latent', lat, lat_o), ('prior', pri, pri_o))}
    ez = (z.detach().double().cpu() - z_o).abs().max().item()
    print('\n[%s] per-example log-lik rel err %s, max |dz| %.3e' % (name, e, ez))
    for k, v in e.items():
        assert v <= TOL_LL[name], (k, v)
    # bits/dim (test_nll.py:61-66)
    bpd = lambda t: (-t.double().mean() / np.log(2.) / spec.ndim).item()
    assert abs(bpd(lp.detach().cpu()) - bpd(lp_o)) <= TOL_LL[name] * abs(bpd(lp_o))
    # hand-written backward vs oracle autograd
    loss_o, grads_o = O.loss_and_grads(spec, P, x)
    loss = cfg.loss(lp)
    params = cfg.model.parameters()
    grads = torch.autograd.grad(loss, params)
    by_id = {id(p): g for p, g in zip(params, grads)}
    worst = (0.0, None)
    # absolute floor: for C = 1 the c1 weight-norm gradient w.r.t. V is exactly 0 in exact arithmetic (W = g sign(V));
    # fp32 leaves the cancellation residue of two O(|g/V| dW) terms there
    floor = 1e-5 * max(v.abs().max().item() for v in grads_o.values())
    for tfname, view in cfg.model.named_tf_variables():
        base = view._base if view._base is not None else view
        g = by_id[id(base)] if id(base) in by_id else by_id[id(view)]
        # re-create the same view on the gradient tensor
        gv = g.as_strided(view.shape, view.stride(), view.storage_offset() - base.storage_offset()).double().cpu()
        ref = grads_o[tfname]
        scale = ref.abs().max().item() + floor
        err = (gv - ref).abs().max().item() / scale
        if err > worst[0]:
            worst = (err, tfname)
        assert err <= TOL_GRAD[name], (tfname, err, scale)
    print('[%s] loss %.6f (oracle %.6f); worst gradient error %.3e of max at %s' % (name, loss.item(), loss_o.item(), worst[0], worst[1]))
    assert abs(loss.item() - loss_o.item()) <= TOL_LL[name] * abs(loss_o.item())
    if name.startswith('en2'):
        return
    # ---- the same comparison against the bf16-emulating oracle: tight tolerance -------------------------------
    with O.emulate_bf16():
        with torch.no_grad():
            lp_e = O.build_model(spec, P, x)[0]
        loss_e, grads_e = O.loss_and_grads(spec, P, x)
    e_emu = rel_err(lp, lp_e)
    worst = (0.0, None)
    floor = 1e-5 * max(v.abs().max().item() for v in grads_e.values())
    for tfname, view in cfg.model.named_tf_variables():
        base = view._base if view._base is not None else view
        g = by_id[id(base)] if id(base) in by_id else by_id[id(view)]
        gv = g.as_strided(view.shape, view.stride(), view.storage_offset() - base.storage_offset()).double().cpu()
        ref = grads_e[tfname]
        err = (gv - ref).abs().max().item() / (ref.abs().max().item() + floor)
        if err > worst[0]:
            worst = (err, tfname)
    print('[%s] vs bf16-emulating oracle: per-example log-lik rel err %.3e, worst gradient error %.3e at %s' % (
        name, e_emu, worst[0], worst[1]))
    assert e_emu <= TOL_LL_EMU
    # an2 with B=1, T=2 on uniform-noise 'bytes' is the extreme regime (|z| ~ 3e4, log p ~ -2e6): gradients chaotic
    # (kernel and emulating oracle are two realisations of that chaos: same bound as against the exact oracle)
    assert worst[0] <= (TOL_GRAD[name] if name.startswith('an2') else TOL_GRAD_EMU), worst


def test_an2_gradients_on_a_non_chaotic_input():
    """an2_cifar_tp (32x32x3, 18 conv couplings) on smooth 'strokes' images, B=1, T=4: unlike the uniform-noise 'bytes' batch of
    the test above (|z| ~ 3e4, gradients chaotic, bound 1.0 = no statement) this input keeps the flow in its ordinary regime,
    so the hand-written backward of the whole conv stack is held to a real bound against the oracle's autograd -- exact
    fp64 oracle and bf16-emulating oracle."""
    name = 'an2_cifar_tp'
    spec, P, x, cfg, xg = setup(name, 1, 4, seed=11, kind='strokes')
    lp = cfg.build_model(xg)[0]
    loss = cfg.loss(lp)
    params = cfg.model.parameters()
    grads = torch.autograd.grad(loss, params)
    by_id = {id(p): g for p, g in zip(params, grads)}

    def worst_against(grads_ref):
        worst = (0.0, None)
        floor = 1e-5 * max(v.abs().max().item() for v in grads_ref.values())
        for tfname, view in cfg.model.named_tf_variables():
            base = view._base if view._base is not None else view
            g = by_id[id(base)] if id(base) in by_id else by_id[id(view)]
            gv = g.as_strided(view.shape, view.stride(), view.storage_offset() - base.storage_offset()).double().cpu()
            ref = grads_ref[tfname]
            err = (gv - ref).abs().max().item() / (ref.abs().max().item() + floor)
            if err > worst[0]:
                worst = (err, tfname)
        return worst
    loss_o, grads_o = O.loss_and_grads(spec, P, x)
    w_exact = worst_against(grads_o)
    with O.emulate_bf16():
        loss_e, grads_e = O.loss_and_grads(spec, P, x)
    w_emu = worst_against(grads_e)
    print('\n[an2 strokes B=1 T=4] loss %.4f (oracle %.4f, emulating %.4f); worst gradient error: %.3e of max at %s (exact oracle), '
          '%.3e at %s (bf16-emulating oracle)' % (loss.item(), loss_o.item(), loss_e.item(), w_exact[0], w_exact[1], w_emu[0], w_emu[1]))
    assert abs(loss.item() - loss_o.item()) <= TOL_LL[name] * abs(loss_o.item())
    assert w_exact[0] <= AN2_TOL_GRAD_EXACT, w_exact
    assert w_emu[0] <= AN2_TOL_GRAD_EMU, w_emu


@pytest.mark.parametrize('name', ['en2_noconv_mnist_tp', 'bn2_omniglot_tp', 'an2_cifar_tp'])
def test_inverse_round_trip(name):
    """backward(forward(x)) == x and the two log-dets cancel (size-independent property of the flow)."""
    spec, P, x, cfg, xg = setup(name, 1, 3, seed=9)
    m = cfg.model
    with torch.no_grad():
        xb = xg.reshape((-1,) + tuple(xg.shape[2:]))
        z, ld = m.flow_forward(xb)
        xr, ldr = m.flow_inverse(z)
    # pixels in [0,256); conv path: the bf16 s/t nets are evaluated twice on slightly different inputs, and mid-range
    # pixels (cifar 'bytes') sit where the sigmoid has slope 64 pixel units per unit of logit
    tol = {'en2_noconv_mnist_tp': 2e-3, 'bn2_omniglot_tp': 0.5, 'an2_cifar_tp': 16.0}[name]
    err = (xr - xb).abs().max().item()
    print('\n[%s] round-trip max |x - x_rec| = %.3e (pixel units), |ld + ld_inv| = %.3e' % (name, err, (ld + ldr).abs().max().item()))
    assert err <= tol
    assert (ld + ldr).abs().max().item() <= (1e-2 if name.startswith('en2') else 2.0)


@pytest.mark.parametrize('name', ['en2_noconv_mnist_tp', 'bn2_omniglot_tp'])
def test_sampling_branch_matches_oracle(name):
    """sampling_mode=True (test_samples.py:48-52) with the random draws supplied: samples from fixed latents."""
    spec, P, x, cfg, xg = setup(name, 1, 3, seed=13)
    T, n, D = 3, cfg.n_samples, spec.ndim
    g = torch.Generator().manual_seed(3)
    noise = [torch.rand(2, n, 1, D, generator=g) for _ in range(T + 1)]
    with torch.no_grad():
        xs_o, lp_o = O.build_model_sampling(spec, P, x, [t.double() for t in noise])
    xs, lp = cfg.build_model(xg, sampling_mode=True, noise=[t.cuda() for t in noise])
    assert tuple(xs.shape) == tuple(xs_o.shape) and tuple(lp.shape) == tuple(lp_o.shape)
    ex = (xs.double().cpu() - xs_o).abs().max().item()
    # The reference's LogitLayer.backward evaluates log1p(exp(y)) (nn_extra_nvp.py:36), which overflows to inf for the
    # large latents an untrained flow produces (column 0 = sample from the prior); the GPU path uses a stable softplus.
    # Compare the log-probs where the reference formula is finite.
    fin = torch.isfinite(lp_o)
    assert torch.isfinite(lp).all()
    el = (((lp.double().cpu() - lp_o).abs() / lp_o.abs().clamp_min(1.0))[fin]).max().item() if fin.any() else 0.0
    print('\n[%s] sampling: max |x_s - oracle| = %.3e (pixel units), log-prob rel err %.3e (%d of %d finite in the oracle)' % (
        name, ex, el, int(fin.sum()), fin.numel()))
    # pixel units in [0, 256); fp32 Bailey sampler + inverse flow (en2), bf16 s/t nets in the inverse pass (bn2)
    assert ex <= (0.5 if name.startswith('en2') else 16.0)
    assert el <= (2e-4 if name.startswith('en2') else 1e-2)


def test_few_shot_argmax_matches_oracle():
    """test_few_shot_omniglot scoring (:64-80): argmax over the k candidate sequences of log_probs[:, -1]."""
    name = 'bn2_omniglot_tp'
    spec = O.spec_for(name)
    P = O.init_params(spec, seed=21, dtype=torch.float64, perturb_process=True)
    from bruno_b200.synthetic_data import SyntheticFewShotIterator
    it = SyntheticFewShotIterator(seq_len=2, batch_size=20, rng=np.random.RandomState(4), n_classes=24)
    x0 = torch.from_numpy(next(it.generate())[0]).double()
    O.data_dependent_init(spec, P, x0)
    P = O.cast_params(O.cast_params(P, torch.float32), torch.float64)
    cfg = fresh_config(name)
    cfg.build_model(x0[:1].float().cuda())
    cfg.model.load_tf_variables(P)
    for ep, (xb, y_batch, x_number) in enumerate(it.generate()):
        if ep >= 3:
            break
        xb = torch.from_numpy(xb)
        with torch.no_grad():
            s_o = O.build_model(spec, P, xb.double())[0][:, -1]
            s_g = cfg.build_model(xb.cuda())[0][:, -1]
        assert int(s_o.argmax()) == int(s_g.argmax())
        top2 = torch.topk(s_o, 2).values
        print('\nfew-shot: argmax %d, oracle margin %.3f, max score err %.3e' % (int(s_g.argmax()), (top2[0] - top2[1]).item(),
                                                                               (s_g.double().cpu() - s_o).abs().max().item()))


def test_conv_coupling_activations_bitwise_vs_emulating_oracle():
    """Layer-level parity of the tensor-core path: c1 output bit-exact, first 3x3 conv output identical except
    < 0.1 % one-ulp (bf16) flips, every slab keeps its zero padding."""
    from bruno_b200 import _lib, slab
    name = 'bn2_omniglot_tp'
    spec, P, x, cfg, xg = setup(name, 2, 3, seed=5)
    layer = cfg.model.nvp_layers[0]
    xb = x.reshape(-1, 28, 28, 1)
    N, H, W, C = xb.shape
    with torch.no_grad():
        y32, _ = O.logit_forward(*O.scale_forward(xb.float(), torch.zeros(N)))
    yin = y32.double()
    acts = []
    with O.emulate_bf16(), torch.no_grad():
        b = O.conv_mask(yin.shape, 'checkerboard0', torch.float64)
        sc = 'model/Checkerboard0_1/function_s_t_wn'
        h = O._act_store(O._wn_conv_pre(yin * b, P, sc + '/c1'))
        acts.append(h)
        u = O._act_store(O._wn_conv_pre(h, P, sc + '/c2_0', quant_w=True))
        acts.append(u)
        acts.append(O._act_store(O._wn_conv_pre(u, P, sc + '/c3_0', quant_w=True) + h))
    R = layer.num_res_blocks
    wf = torch.empty(2 * R, _lib.query('bruno_conv_wpack_bytes'), dtype=torch.uint8, device='cuda')
    _lib.call('bruno_conv_wn_pack', layer.Vr, layer.gr, wf, None, 2 * R)
    slabs = slab.alloc_slabs(2 * R + 1, N, H, W, 'cuda')
    yg = yin.float().cuda()
    ld = torch.zeros(N, device='cuda')
    _lib.call('bruno_coupling_conv_fwd', yg, ld, torch.empty_like(yg), torch.empty_like(ld), N, H, W, C, R, layer.mask_id, 0,
              layer.V1, layer.g1, layer.b1, wf, layer.br, 0, layer.Ks, layer.bs, layer.Kt, layer.bt, None, None, slabs, 2 * R + 1, 1)
    frac = []
    for i, a in enumerate(acts):
        got = slab.slab_to_nhwc(slabs[i], N, H, W, check_padding=True).double().cpu()
        frac.append((got != a).double().mean().item())
        ulp = (got - a).abs().max().item() / max(a.abs().max().item(), 1e-9)
        assert ulp <= 2.0 ** -7
    print('\nmismatching bf16 values: c1 %.4f%%, c2_0 %.4f%%, block-0 output %.4f%%' % tuple(100 * f for f in frac))
    assert frac[0] == 0.0 and frac[1] < 1e-3 and frac[2] < 2e-3


@pytest.mark.parametrize('name,B,T', [('bn2_omniglot_tp', 2, 3), ('bn2_omniglot_gp', 2, 2), ('an2_cifar_tp', 1, 2)])
def test_fp32_verification_mode_meets_the_fp32_bound(name, B, T):
    """The conv configs evaluated with nn_extra_nvp.precision('fp32') (every convolution in fp32 on CUDA cores) meet the
    north star's fp32 bound -- per-example log-likelihood within 1e-4 relative of the oracle -- so the gap of the
    production path is the declared bf16 rounding and nothing else."""
    from bruno_b200 import nn_extra_nvp
    spec, P, x, cfg, xg = setup(name, B, T, seed=5)
    with torch.no_grad():
        lp_o, lat_o, pri_o, z_o = O.build_model(spec, P, x)
        with nn_extra_nvp.precision('fp32'):
            lp, lat, pri, z = cfg.build_model(xg)
        lp16 = cfg.build_model(xg)[0]
    e32, e16 = rel_err(lp, lp_o), rel_err(lp16, lp_o)
    zmax = z_o.abs().max().item()
    print('\n[%s] per-example log-lik rel err: fp32 verification mode %.3e, bf16 production mode %.3e; max |dz| fp32 %.3e (|z| max %.1f)' % (
        name, e32, e16, (z.double().cpu() - z_o).abs().max().item(), zmax))
    assert e32 <= 1e-4
    assert rel_err(lat, lat_o) <= 1e-4 and rel_err(pri, pri_o) <= 1e-4


def test_few_shot_argmax_fp32_and_bf16_agree_with_oracle():
    """Bit-exact few-shot decisions (north star): argmax over the 20 candidates in fp32 verification mode and in the
    bf16 production mode both equal the oracle's, on every tested episode."""
    from bruno_b200 import nn_extra_nvp
    from bruno_b200.synthetic_data import SyntheticFewShotIterator
    name = 'bn2_omniglot_tp'
    spec = O.spec_for(name)
    P = O.init_params(spec, seed=33, dtype=torch.float64, perturb_process=True)
    it = SyntheticFewShotIterator(seq_len=2, batch_size=20, rng=np.random.RandomState(11), n_classes=24)
    eps = [e for e, _ in zip(it.generate(), range(4))]
    O.data_dependent_init(spec, P, torch.from_numpy(eps[0][0]).double())
    P = O.cast_params(O.cast_params(P, torch.float32), torch.float64)
    cfg = fresh_config(name)
    cfg.build_model(torch.from_numpy(eps[0][0][:1]).cuda())
    cfg.model.load_tf_variables(P)
    for xb, y_batch, x_number in eps:
        xb = torch.from_numpy(xb)
        with torch.no_grad():
            s_o = O.build_model(spec, P, xb.double())[0][:, -1]
            s16 = cfg.build_model(xb.cuda())[0][:, -1]
            with nn_extra_nvp.precision('fp32'):
                s32 = cfg.build_model(xb.cuda())[0][:, -1]
        assert int(s_o.argmax()) == int(s32.argmax()) == int(s16.argmax())
        assert rel_err(s32, s_o) <= 1e-4


def test_cuda_graphed_forward_equals_eager():
    """The CUDA-graphed few-shot forward (one graph launch per episode) is bit-identical to the eager path."""
    name = 'bn2_omniglot_tp'
    spec, P, x, cfg, xg = setup(name, 3, 2, seed=17)
    with torch.no_grad():
        ref = cfg.build_model(xg, want_prior=False)[0].clone()
        f = cfg.model.graphed_forward(xg)
        out = f(xg)[0].clone()
        assert torch.equal(out, ref)
        x2 = xg.flip(0).contiguous()
        ref2 = cfg.build_model(x2, want_prior=False)[0].clone()
        assert torch.equal(f(x2)[0], ref2)
