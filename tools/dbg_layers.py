"""Per-layer divergence of the GPU flow from the oracle (exact and bf16-emulating) + bitwise conv check."""
import sys, importlib, torch, numpy as np
sys.path.insert(0, '/root/repo')
from oracle import bruno_oracle as O
import torch.nn.functional as F
from bruno_b200 import _lib, slab, nn_extra_nvp

name = 'bn2_omniglot_tp'
spec = O.spec_for(name)
P = O.init_params(spec, seed=5, dtype=torch.float64, perturb_process=True)
x = O.synthetic_batch(spec, 2, 3, seed=6).float().double()
O.data_dependent_init(spec, P, x)
P = O.cast_params(O.cast_params(P, torch.float32), torch.float64)
cfg = importlib.import_module('bruno_b200.config_rnn.' + name)
xg = x.float().cuda()
cfg.build_model(xg[:1]); cfg.model.load_tf_variables(P)
xb = x.reshape(-1, 28, 28, 1)
tr, tre = {}, {}
with torch.no_grad():
    O.flow_forward(spec, P, xb, trace=tr)
    with O.emulate_bf16():
        O.flow_forward(spec, P, xb, trace=tre)
m = cfg.model
with torch.no_grad():
    y, ld = nn_extra_nvp.preprocess_forward(xg.reshape(-1, 28, 28, 1))
    print('pre: y err %.3e ld err %.3e' % ((y.double().cpu() - tr['pre'][0]).abs().max(), (ld.double().cpu() - tr['pre'][1]).abs().max()))
    z = None
    for layer in m.nvp_layers:
        y, z, ld = layer.forward_and_jacobian(y, z, ld)
        a, b = tr[layer.name], tre[layer.name]
        print('%-16s |y|max %.2f  y err exact %.3e  emu %.3e   ld err exact %.3e emu %.3e' % (
            layer.name, a[0].abs().max(), (y.double().cpu() - a[0]).abs().max(), (y.double().cpu() - b[0]).abs().max(),
            (ld.double().cpu() - a[2]).abs().max(), (ld.double().cpu() - b[2]).abs().max()))

# bitwise conv check
def bf(t): return t.to(torch.bfloat16).to(torch.float64)
g = torch.Generator().manual_seed(0)
N, H, W = 4, 28, 28
V = torch.randn(1, 3, 3, 64, 64, generator=g, dtype=torch.float64) * 0.05
gain = torch.ones(1, 64, dtype=torch.float64) * 3
bias = torch.randn(64, generator=g, dtype=torch.float64) * 0.2
xx = torch.randn(N, H, W, 64, generator=g, dtype=torch.float64)
wb = _lib.query('bruno_conv_wpack_bytes')
wf = torch.empty(1, wb, dtype=torch.uint8, device='cuda')
_lib.call('bruno_conv_wn_pack', V.float().cuda().contiguous(), gain.float().cuda().contiguous(), wf, None, 1)
nrm = torch.sqrt((V[0].float().double() ** 2).sum(dim=(0, 1, 2), keepdim=True))
Wt = bf(gain[0].float().double().view(1, 1, 1, -1) * V[0].float().double() / nrm)
xs = slab.nhwc_to_slab(xx.float().cuda())
out = slab.alloc_slabs(1, N, H, W, 'cuda')[0]
_lib.call('bruno_conv3x3', 4, xs, wf[0], bias.float().cuda(), 0, None, None, out, N, H, W)
got = slab.slab_to_nhwc(out, N, H, W).double().cpu()
ref64 = F.conv2d(bf(xx).permute(0, 3, 1, 2), Wt.permute(3, 2, 0, 1), padding=1).permute(0, 2, 3, 1) + bias.float().double()
ref = bf(ref64)
mism = (got != ref)
print('conv linear: mismatching bf16 outputs %.4f%%, max |diff|/|ref| %.3e, fp32-level err of acc: n/a' % (100 * mism.double().mean(), ((got - ref).abs() / ref.abs().clamp_min(1e-3)).max()))
# how far is the fp32 accumulator from exact?  use a tiny-magnitude probe: compare against fp64 before rounding
print('   mean |got - ref64| / mean|ref64| = %.3e (bf16 rounding alone gives ~1e-3)' % ((got - ref64).abs().mean() / ref64.abs().mean()))
