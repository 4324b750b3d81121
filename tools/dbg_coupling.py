"""Activation-by-activation comparison of ONE conv coupling (GPU slabs vs bf16-emulating oracle)."""
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
layer = cfg.model.nvp_layers[0]
xb = x.reshape(-1, 28, 28, 1)
N, H, W, C = xb.shape
with torch.no_grad():
    y32, ld32 = O.logit_forward(*O.scale_forward(xb.float(), torch.zeros(N)))
yin = y32.double()
# oracle, emulated, step by step
O.EMULATE_BF16 = True
acts = []
with torch.no_grad():
    b = O.conv_mask(yin.shape, 'checkerboard0', torch.float64)
    sc = 'model/Checkerboard0_1/function_s_t_wn'
    h = O._act_store(O._wn_conv_pre(yin * b, P, sc + '/c1'))
    acts.append(h)
    skip = h
    for r in range(8):
        u = O._act_store(O._wn_conv_pre(h, P, sc + '/c2_%d' % r, quant_w=True)); acts.append(u)
        h = O._act_store(O._wn_conv_pre(u, P, sc + '/c3_%d' % r, quant_w=True) + skip); acts.append(h)
        skip = h
# gpu
yg = yin.float().cuda()
R = 8
wb = _lib.query('bruno_conv_wpack_bytes')
wf = torch.empty(2 * R, wb, dtype=torch.uint8, device='cuda')
_lib.call('bruno_conv_wn_pack', layer.Vr, layer.gr, wf, None, 2 * R)
slabs = slab.alloc_slabs(2 * R + 1, N, H, W, 'cuda')
ld = torch.zeros(N, device='cuda'); out = torch.empty_like(yg); ldo = torch.empty_like(ld)
_lib.call('bruno_coupling_conv_fwd', yg, ld, out, ldo, N, H, W, C, R, layer.mask_id, 0, layer.V1, layer.g1, layer.b1, wf,
          layer.br, 0, layer.Ks, layer.bs, layer.Kt, layer.bt, None, None, slabs, 2 * R + 1, 1)
for i, a in enumerate(acts):
    got = slab.slab_to_nhwc(slabs[i], N, H, W, check_padding=True).double().cpu()
    mism = (got != a).double().mean().item()
    print('act %2d  mismatch %.4f%%  max abs diff %.3e  mean abs diff %.3e  (|a| mean %.3f max %.2f)' % (
        i, 100 * mism, (got - a).abs().max(), (got - a).abs().mean(), a.abs().mean(), a.abs().max()))
