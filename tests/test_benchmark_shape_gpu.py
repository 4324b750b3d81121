"""GPU parity AT THE SHAPES bench.py MEASURES (BASELINE.json configs C2, C3, C4), not only at toy sizes:

  * bn2_omniglot_tp B=32, T=20 (640 images; ~4200 conv tiles at 28x28 = ~28 per persistent CTA) forward NLL,
  * the few-shot shape B=20, T=2, and an2_cifar_tp B=16, T=20 forward,
  * gradients of bn2_omniglot_tp at B=8, T=20 (160 images: > 1 tile per CTA in every conv launch, and small enough that
    the float64 oracle's autograd graph fits in HBM),

each against the float64 oracle (pinned to the reference's code, tests/test_reference_pinning_cpu.py) evaluated on the
GPU through torch so that the full shape takes seconds (tests/_gpu_util.oracle_device).  Bounds: the stated bf16 bound
6e-3 relative per-example log-likelihood in production mode, the north star's fp32 bound 1e-4 in
nn_extra_nvp.precision('fp32').  Reference sites: config_rnn/bn2_omniglot_tp.py:10,59-168; an2_cifar_tp.py:10,44-153;
config_rnn/test_few_shot_omniglot.py:64-80."""
import numpy as np
import pytest
import torch

from oracle import bruno_oracle as O
from _gpu_util import load_model, rel_err, oracle_forward, oracle_device, params_to, grads_by_tf_name

pytestmark = pytest.mark.gpu


def _case(name, B, T, seed, kind):
    spec = O.spec_for(name)
    P = O.init_params(spec, seed=seed, dtype=torch.float64, perturb_process=True)
    x = O.synthetic_batch(spec, B, T, seed=seed + 1, kind=kind).float().double()
    # data-dependent init on a small slice (CPU oracle), like train.py's iteration 0
    O.data_dependent_init(spec, P, x[:2, :min(T, 4)])
    P = O.cast_params(O.cast_params(P, torch.float32), torch.float64)
    return spec, P, x


@pytest.mark.parametrize('name,B,T,kind', [('bn2_omniglot_tp', 32, 20, 'strokes'), ('bn2_omniglot_tp', 20, 2, 'strokes'),
                                           ('bn2_omniglot_gp', 32, 20, 'strokes'), ('an2_cifar_tp', 16, 20, 'bytes')])
def test_forward_nll_at_benchmark_shape(name, B, T, kind):
    from bruno_b200 import nn_extra_nvp
    spec, P, x = _case(name, B, T, seed=41, kind=kind)
    lp_o, lat_o, pri_o, z_o = oracle_forward(spec, P, x)
    cfg = load_model(name, P, x)
    xg = x.float().cuda()
    with torch.no_grad():
        lp, lat, pri, z = cfg.build_model(xg)
        with nn_extra_nvp.precision('fp32'):
            lp32, lat32, pri32, z32 = cfg.build_model(xg)
    e16 = {k: rel_err(a, b) for k, a, b in (('log_probs', lp, lp_o), ('latent', lat, lat_o), ('prior', pri, pri_o))}
    e32 = {k: rel_err(a, b) for k, a, b in (('log_probs', lp32, lp_o), ('latent', lat32, lat_o), ('prior', pri32, pri_o))}
    bpd = lambda t: (-t.double().mean() / np.log(2.) / spec.ndim).item()          # test_nll.py:61-66
    print('\n[%s B=%d T=%d] per-example log-lik rel err: bf16 %s | fp32 mode %s | bits/dim %.6f (oracle %.6f)' % (
        name, B, T, {k: '%.2e' % v for k, v in e16.items()}, {k: '%.2e' % v for k, v in e32.items()}, bpd(lp.cpu()), bpd(lp_o)))
    # Stated bf16 bound: 6e-3 for the Student-t process.  The GAUSSIAN process at this untrained, perturbed-parameter
    # operating point (|log p| ~ 5e6 nats per frame, bits/dim ~ 9e3 instead of the ~1-3 of a trained model) is quadratic
    # in z: rel err(log p) = 2 x rel err(z), and the posterior variance shrinks with every observation, so the same
    # ~4e-3 bf16 error of z that the t-process damps through log1p shows up doubled: stated bound 2e-2 (measured 1.2e-2).
    tol16 = 2e-2 if name.endswith('_gp') else 6e-3
    for k in e16:
        assert e16[k] <= tol16, (k, e16[k])
        assert e32[k] <= 1e-4, (k, e32[k])
    assert abs(bpd(lp.cpu()) - bpd(lp_o)) <= tol16 * abs(bpd(lp_o))
    assert abs(bpd(lp32.cpu()) - bpd(lp_o)) <= 1e-4 * abs(bpd(lp_o))
    # every sequence is checked on its own: a mis-handled tile corrupts a few images, not the mean
    per_seq = ((lp.double().cpu() - lp_o).abs() / lp_o.abs().clamp_min(1.0)).max(dim=1).values
    assert int((per_seq > tol16).sum()) == 0


def test_few_shot_argmax_at_reference_shape():
    """20-way 1-shot episodes (B=20, T=2): the argmax over candidates of log_probs[:, -1] equals the float64 oracle's in
    production (bf16) and fp32 mode on every episode (test_few_shot_omniglot.py:64-80)."""
    from bruno_b200 import nn_extra_nvp
    from bruno_b200.synthetic_data import SyntheticFewShotIterator
    name = 'bn2_omniglot_tp'
    spec = O.spec_for(name)
    P = O.init_params(spec, seed=43, dtype=torch.float64, perturb_process=True)
    it = SyntheticFewShotIterator(seq_len=2, batch_size=20, rng=np.random.RandomState(7), n_classes=30)
    eps = [e for e, _ in zip(it.generate(), range(8))]
    O.data_dependent_init(spec, P, torch.from_numpy(eps[0][0][:4]).double())
    P = O.cast_params(O.cast_params(P, torch.float32), torch.float64)
    cfg = load_model(name, P, torch.from_numpy(eps[0][0]))
    for xb, y_batch, x_number in eps:
        xb = torch.from_numpy(xb)
        s_o = oracle_forward(spec, P, xb.double())[0][:, -1]
        with torch.no_grad():
            s16 = cfg.build_model(xb.cuda())[0][:, -1]
            with nn_extra_nvp.precision('fp32'):
                s32 = cfg.build_model(xb.cuda())[0][:, -1]
        assert int(s_o.argmax()) == int(s32.argmax()) == int(s16.argmax())
        assert rel_err(s32, s_o) <= 1e-4


def test_gradients_with_many_tiles_per_cta():
    """bn2_omniglot_tp B=8, T=20 = 160 images: 1051 / 282 conv tiles at 28x28 / 14x14 (up to 8 per CTA) in forward, dgrad
    and wgrad.  The hand-written backward is compared with autograd of the float64 oracle (run on the GPU) for every
    variable: production (bf16) mode within the stated 1e-1 of each tensor's max, and the loss within 6e-3."""
    name, B, T = 'bn2_omniglot_tp', 8, 20
    spec, P, x = _case(name, B, T, seed=47, kind='strokes')
    with oracle_device('cuda'):
        loss_o, grads_o = O.loss_and_grads(spec, params_to(P, 'cuda'), x.cuda())
    grads_o = {k: v.cpu() for k, v in grads_o.items()}
    torch.cuda.empty_cache()
    cfg = load_model(name, P, x)
    lp = cfg.build_model(x.float().cuda())[0]
    loss = cfg.loss(lp)
    grads = grads_by_tf_name(cfg, loss)
    gmax = max(v.abs().max().item() for v in grads_o.values())
    worst = (0.0, None)
    ranked = []
    for k, ref in grads_o.items():
        # a c1 with ONE input channel has W = g sign(V): d loss / d V is exactly 0 and what either side holds is the
        # cancellation residue of two O(|g / V| dW) terms -- judged against the model's largest gradient, not its own max
        floor = (1e-3 if (k.endswith('/c1/V') and ref.shape[2] == 1) else 1e-5) * gmax
        err = (grads[k] - ref).abs().max().item() / (ref.abs().max().item() + floor)
        ranked.append((err, k))
        if err > worst[0]:
            worst = (err, k)
    print('\n worst five: %s' % sorted(ranked, reverse=True)[:5])
    print('\n[%s B=%d T=%d] loss %.6f (oracle %.6f), worst gradient error %.3e of max at %s' % (
        name, B, T, loss.item(), loss_o.item(), worst[0], worst[1]))
    assert abs(loss.item() - loss_o.item()) <= 6e-3 * abs(loss_o.item())
    assert worst[0] <= 1e-1, worst
