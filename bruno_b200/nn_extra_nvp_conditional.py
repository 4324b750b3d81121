"""Label-conditioned Real NVP layers -- host mirror of the reference's nn_extra_nvp_conditional.py (Conditional BRUNO,
config_conditional/m1_shapenet.py).  Layer protocol of the reference (:27-31):
    forward_and_jacobian(x, sum_log_det_jacobians, z, y_label) -> (y, sum_log_det_jacobians, z)
    backward(y, z, y_label) -> (x, z)
Every weight-normalised conv adds dense_wn(y_label)[:, None, None, :] to its output (:422-433) -- on the GPU that is a
per-image bias table consumed by the tensor-core conv epilogue; every dense layer is fed concat(dense(y_label), x)
(:461-471, :350-369).  Round 1 implements the INFERENCE path (NLL evaluation and sampling, forward + inverse); the
hand-written backward of the conditional layers is not built yet and training raises NotImplementedError.
"""
import numpy as np
import torch

from . import _lib
from .nn_extra_nvp import (MASKS, SLABS, Layer as _BaseLayer, elu, int_shape, _new_like, _pre, _SpaceToDepthFn,  # noqa: F401
                           channel_slice, channel_concat)

LEAKY, TANH, ELU, NONE = 2, 3, 1, 0


def logit_forward_and_jacobian(x, sum_log_det_jacobians):                  # cond:10-16 (no ln(1 - alpha) term)
    xs = int_shape(x)
    y = torch.empty_like(x)
    ld = _new_like(x, xs[0])
    _lib.call('bruno_pre', 1, x.contiguous(), sum_log_det_jacobians.contiguous(), y, ld, xs[0], int(np.prod(xs[1:])), 256., 1e-5, 1)
    return y, ld


def dequantization_forward_and_jacobian(x, sum_log_det_jacobians):         # cond:19-23
    xs = int_shape(x)
    y = torch.empty_like(x)
    ld = _new_like(x, xs[0])
    _lib.call('bruno_pre', 0, x.contiguous(), sum_log_det_jacobians.contiguous(), y, ld, xs[0], int(np.prod(xs[1:])), 256., 1e-5, 1)
    return y, ld


def _sgemm(a, b, out_cols, col_scale=None, bias=None, act=NONE):
    """act((a @ b) * col_scale + bias) with a [M,K], b [K,N] contiguous fp32."""
    M, K = a.shape
    c = _new_like(a, M, out_cols)
    _lib.call('bruno_sgemm', 0, 0, M, out_cols, K, a.contiguous(), K, b.contiguous(), out_cols, c, out_cols, col_scale, bias, act, 0)
    return c


def dense_wn_apply(x, V, g, b, act=NONE, extra_bias=None):
    """dense_wn (cond:437-459): (x @ V) * g / ||V|| + b."""
    K, N = V.shape
    sc = _new_like(x, N)
    _lib.call('bruno_dense_wn_scale', V.contiguous(), g.contiguous(), sc, None, K, N)
    bias = b if extra_bias is None else (b + extra_bias).contiguous()
    return _sgemm(x, V, N, sc, bias.contiguous(), act)


class Layer(object):                                                        # cond:26-31
    def forward_and_jacobian(self, x, sum_log_det_jacobians, z, y_label):
        raise NotImplementedError(str(type(self)))

    def backward(self, y, z, y_label):
        raise NotImplementedError(str(type(self)))

    def named_tf_variables(self, prefix='model/'):
        return []


def _normal(*s, device, std=0.05):
    return torch.randn(*s, device=device) * std


class CouplingLayerConv(Layer):                                             # cond:34-159
    def __init__(self, mask_type, name='CouplingLayer', nonlinearity=elu, weight_norm=True, num_filters=64, num_res_blocks=8):
        assert mask_type in ('checkerboard0', 'checkerboard1', 'channel0', 'channel1')
        if not weight_norm or num_filters != 64 or nonlinearity is not elu:
            raise NotImplementedError('B200 path: weight-normalised, 64 filters, ELU')
        self.mask_type, self.mask_id, self.name = mask_type, MASKS[mask_type], name
        self.nonlinearity, self.weight_norm = nonlinearity, weight_norm
        self.num_filters, self.num_res_blocks = num_filters, num_res_blocks
        self.V1 = None

    def _build(self, C, L, device):
        F, R = self.num_filters, self.num_res_blocks
        nl = 2 * R + 1
        self.V1 = _normal(C, F, device=device)
        self.g1, self.b1 = torch.ones(F, device=device), torch.zeros(F, device=device)
        self.Vr = _normal(2 * R, 3, 3, F, F, device=device)
        self.gr, self.br = torch.ones(2 * R, F, device=device), torch.zeros(2 * R, F, device=device)
        # label paths of the 2R+1 weight-normalised convs, concatenated along the output dim: [L, (2R+1)*64]
        self.LV = _normal(L, nl * F, device=device)
        self.Lg, self.Lb = torch.ones(nl * F, device=device), torch.zeros(nl * F, device=device)
        self.Ks, self.bs = torch.zeros(F, C, device=device), torch.zeros(C, device=device)
        self.Kt, self.bt = torch.zeros(F, C, device=device), torch.zeros(C, device=device)
        self.LKs, self.LKt = _normal(L, C, device=device, std=0.3), _normal(L, C, device=device, std=0.3)

    def named_tf_variables(self, prefix='model/'):
        sc = '%s%s/function_s_t_wn/' % (prefix, self.name)
        F, R, C = self.num_filters, self.num_res_blocks, self.V1.shape[0]
        names = ['c1'] + [n for r in range(R) for n in ('c2_%d' % r, 'c3_%d' % r)]
        out = []
        for i, nm in enumerate(names):
            out += [(sc + nm + '/label_h/V', self.LV[:, i * F:(i + 1) * F]), (sc + nm + '/label_h/g', self.Lg[i * F:(i + 1) * F]),
                    (sc + nm + '/label_h/b', self.Lb[i * F:(i + 1) * F])]
            if i == 0:
                out += [(sc + 'c1/conv_h/V', self.V1.view(1, 1, C, F)), (sc + 'c1/conv_h/g', self.g1), (sc + 'c1/conv_h/b', self.b1)]
            else:
                out += [(sc + nm + '/conv_h/V', self.Vr[i - 1]), (sc + nm + '/conv_h/g', self.gr[i - 1]),
                        (sc + nm + '/conv_h/b', self.br[i - 1])]
        for nm, K, b, LK in (('conv_out_scale', self.Ks, self.bs, self.LKs), ('conv_out_translation', self.Kt, self.bt, self.LKt)):
            out += [('%s%s/label_h/kernel' % (sc, nm), LK), ('%s%s/%s/kernel' % (sc, nm, nm), K.view(1, 1, F, C)),
                    ('%s%s/%s/bias' % (sc, nm, nm), b)]
        return out

    def _apply(self, x, ld, y_label, inverse):
        if torch.is_grad_enabled() and (x.requires_grad or (ld is not None and ld.requires_grad)):
            raise NotImplementedError('conditional layers: backward (training) is not built in round 1')
        x = x.contiguous()
        N, H, W, C = x.shape
        if self.V1 is None:
            self._build(C, int(y_label.shape[1]), x.device)
        F, R = self.num_filters, self.num_res_blocks
        nl = 2 * R + 1
        with torch.no_grad():
            # per-image bias tables [N, (2R+1)*64]: dense_wn(y_label) + b of the conv (cond:422-433)
            conv_b = torch.cat([self.b1.view(1, F), self.br], 0).reshape(-1)
            tab = dense_wn_apply(y_label, self.LV, self.Lg, self.Lb, NONE, extra_bias=conv_b)
            ls = _sgemm(y_label, self.LKs, C)
            lt = _sgemm(y_label, self.LKt, C)
            wfwd = torch.empty(2 * R, _lib.query('bruno_conv_wpack_bytes'), dtype=torch.uint8, device=x.device)
            _lib.call('bruno_conv_wn_pack', self.Vr, self.gr, wfwd, None, 2 * R)
            slabs = SLABS.get('pingpong', 3, N, H, W, x.device)
            y = torch.empty_like(x)
            ld_in = ld.contiguous() if ld is not None else torch.zeros(N, device=x.device)
            ld_out = torch.empty_like(ld_in)
            _lib.call('bruno_coupling_conv_fwd', x, ld_in, y, ld_out, N, H, W, C, R, self.mask_id, int(inverse), self.V1,
                      self.g1, tab, wfwd, tab.data_ptr() + 4 * F, nl * F, self.Ks, self.bs, self.Kt, self.bt, ls, lt, slabs, 3)
        return y, ld_out

    def forward_and_jacobian(self, x, sum_log_det_jacobians, z, y_label=None):      # cond:137-149
        y, ld = self._apply(x, sum_log_det_jacobians, y_label, False)
        return y, ld, z

    def backward(self, y, z, y_label=None):                                         # cond:151-159 (no log-det)
        x, _ = self._apply(y, None, y_label, True)
        return x, z


class CouplingLayerDense(Layer):                                            # cond:162-266
    def __init__(self, mask_type, name='CouplingDense', nonlinearity=elu, n_units=1024, weight_norm=True):
        assert mask_type in ('even', 'odd')
        if not weight_norm or nonlinearity is not elu:
            raise NotImplementedError('B200 path: weight-normalised, ELU')
        self.mask_type, self.mask_id, self.name = mask_type, MASKS[mask_type], name
        self.nonlinearity, self.weight_norm, self.n_units = nonlinearity, weight_norm, n_units
        self.P = None

    def _build(self, D, L, device):
        U = self.n_units
        P = {}
        for nm, din in (('d1', D), ('d2', U)):
            P[nm + '/label_h/V'] = _normal(L, din // 2, device=device)
            P[nm + '/label_h/g'] = torch.ones(din // 2, device=device)
            P[nm + '/label_h/b'] = torch.zeros(din // 2, device=device)
            P['%s/%s/V' % (nm, nm)] = _normal(din // 2 + din, U, device=device)
            P['%s/%s/g' % (nm, nm)] = torch.ones(U, device=device)
            P['%s/%s/b' % (nm, nm)] = torch.zeros(U, device=device)
        for nm in ('d_scale', 'd_translate'):
            P[nm + '/label_W1/kernel'] = _normal(L, U // 2, device=device, std=0.3)
            P[nm + '/label_W1/bias'] = torch.zeros(U // 2, device=device)
            P['%s/%s/kernel' % (nm, nm)] = torch.zeros(U // 2 + U, D, device=device)
            P['%s/%s/bias' % (nm, nm)] = torch.zeros(D, device=device)
        self.P = P

    def named_tf_variables(self, prefix='model/'):
        sc = '%s%s/function_s_t_dense_wn/' % (prefix, self.name)
        return [(sc + k, v) for k, v in self.P.items()]

    def _apply(self, x, ld, y_label, inverse):
        if torch.is_grad_enabled() and (x.requires_grad or (ld is not None and ld.requires_grad)):
            raise NotImplementedError('conditional layers: backward (training) is not built in round 1')
        xs = x.shape
        x = x.contiguous()
        N, D = xs[0], int(np.prod(xs[1:]))
        if self.P is None:
            self._build(D, int(y_label.shape[1]), x.device)
        P = self.P
        with torch.no_grad():
            idx = torch.arange(D, device=x.device) % 2
            b = (idx if self.mask_type == 'even' else 1 - idx).to(torch.float32)
            y = x.view(N, D) * b                                                       # x1 = x * mask
            for nm in ('d1', 'd2'):                                                    # cond:461-471
                h = dense_wn_apply(y_label, P[nm + '/label_h/V'], P[nm + '/label_h/g'], P[nm + '/label_h/b'], LEAKY)
                y = dense_wn_apply(torch.cat([h, y], 1), P['%s/%s/V' % (nm, nm)], P['%s/%s/g' % (nm, nm)],
                                   P['%s/%s/b' % (nm, nm)], ELU)
            pre = []
            for nm in ('d_scale', 'd_translate'):                                      # cond:350-369
                h = _sgemm(y_label, P[nm + '/label_W1/kernel'], self.n_units // 2, None, P[nm + '/label_W1/bias'], LEAKY)
                pre.append(_sgemm(torch.cat([h, y], 1), P['%s/%s/kernel' % (nm, nm)], D, None, P['%s/%s/bias' % (nm, nm)], NONE))
            out = torch.empty_like(x)
            ld_in = ld.contiguous() if ld is not None else torch.zeros(N, device=x.device)
            ld_out = torch.empty_like(ld_in)
            _lib.call('bruno_dense_coupling', x, pre[0], pre[1], ld_in, out, ld_out, N, D, self.mask_id, int(inverse))
        return out, ld_out

    def forward_and_jacobian(self, x, sum_log_det_jacobians, z, y_label=None):
        y, ld = self._apply(x, sum_log_det_jacobians, y_label, False)
        return y, ld, z

    def backward(self, y, z, y_label=None):
        x, _ = self._apply(y, None, y_label, True)
        return x, z


class SqueezingLayer(Layer):                                                # cond:269-291
    def __init__(self, name="Squeeze"):
        self.name = name

    def forward_and_jacobian(self, x, sum_log_det_jacobians, z, y_label=None):
        xs = int_shape(x)
        assert xs[1] % 2 == 0 and xs[2] % 2 == 0
        y = _SpaceToDepthFn.apply(x, False)
        if z is not None:
            z = _SpaceToDepthFn.apply(z, False)
        return y, sum_log_det_jacobians, z

    def backward(self, y, z, y_label=None):
        assert int_shape(y)[3] % 4 == 0
        x = _SpaceToDepthFn.apply(y, True)
        if z is not None:
            z = _SpaceToDepthFn.apply(z, True)
        return x, z


class FactorOutLayer(Layer):                                                # cond:294-338
    def __init__(self, scale, name='FactorOut'):
        self.scale, self.name = scale, name

    def forward_and_jacobian(self, x, sum_log_det_jacobians, z, y_label=None):
        xs = int_shape(x)
        split = xs[3] // 2
        new_z = channel_slice(x, 0, split)
        x = channel_slice(x, split, xs[3] - split)
        z = channel_concat(z, new_z) if z is not None else new_z
        return x, sum_log_det_jacobians, z

    def backward(self, y, z, y_label=None):
        zs = int_shape(z)
        split = zs[3] // (2 ** self.scale) if y is None else int_shape(y)[3]
        new_y = channel_slice(z, zs[3] - split, split)
        z = channel_slice(z, 0, zs[3] - split) if zs[3] - split > 0 else None
        assert int_shape(new_y)[3] == split
        x = channel_concat(new_y, y) if y is not None else new_y
        return x, z
