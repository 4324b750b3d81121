"""Label-conditioned Real NVP layers -- host mirror of the reference's nn_extra_nvp_conditional.py (Conditional BRUNO,
config_conditional/m1_shapenet.py).  Layer protocol of the reference (:27-31):
    forward_and_jacobian(x, sum_log_det_jacobians, z, y_label) -> (y, sum_log_det_jacobians, z)
    backward(y, z, y_label) -> (x, z)
Every weight-normalised conv adds dense_wn(y_label)[:, None, None, :] to its output (:422-433) -- on the GPU that is a
per-image bias table consumed by the tensor-core conv epilogue; every dense layer is fed concat(dense(y_label), x)
(:461-471, :350-369).
Inference (NLL evaluation, sampling: forward + inverse) and training run through the fused kernels; every backward is
hand-written: the conv coupling (bruno_coupling_conv_bwd_cond: the unconditional backward + per-image gradients of the
bias tables and head terms), the dense coupling (bruno_coupling_dense_cond_bwd: one call for all 20 variables, x and the
label embedding) and the label paths (bruno_dense_layer_bwd per dense layer).  torch only carries the autograd graph
between those calls.
"""
import numpy as np
import torch

from . import _lib
from .nn_extra_nvp import (MASKS, SLABS, _check_generation, Layer as _BaseLayer, elu, int_shape, _new_like, _pre, _SpaceToDepthFn,  # noqa: F401
                           channel_slice, channel_concat)
from .slab import alloc_slabs

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


# ------------------------------------------------------------------------------------------------------------------
# differentiable primitives of the training path
# ------------------------------------------------------------------------------------------------------------------
class _DenseLayerFn(torch.autograd.Function):
    """One dense layer of a label path -- dense_wn (g given) or plain dense -- with fused scale / bias / activation; the
    backward (dx, dW or dV, dg, db) is the hand-written bruno_dense_layer_bwd, not autograd of torch ops."""
    @staticmethod
    def forward(ctx, x, W, g, b, extra_bias, act):
        x, W = x.contiguous(), W.contiguous()
        N, K = x.shape
        U = W.shape[1]
        out, pre = _new_like(x, N, U), _new_like(x, N, U)
        sc = _new_like(x, U) if g is not None else None
        inv = _new_like(x, U) if g is not None else None
        btot = _new_like(x, U)
        _lib.call('bruno_dense_layer_fwd', x, W, g, b, extra_bias, N, K, U, int(act), out, pre, sc, inv, btot)
        ctx.act, ctx.has = int(act), (g is not None, b is not None, extra_bias is not None)
        ctx.save_for_backward(x, W, g, btot, sc, inv, pre, out)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, W, g, btot, sc, inv, pre, out = ctx.saved_tensors
        N, K = x.shape
        U = W.shape[1]
        ws = _new_like(x, _lib.query('bruno_dense_layer_bwd_ws_floats', N, U))
        dx = _new_like(x, N, K) if ctx.needs_input_grad[0] else None
        dW, db = torch.empty_like(W), _new_like(x, U)
        dg = _new_like(x, U) if g is not None else None
        _lib.call('bruno_dense_layer_bwd', x, W, g, btot, sc, inv, pre, out, dout.contiguous(), N, K, U, ctx.act, ws, dx, dW, dg, db)
        has_g, has_b, has_e = ctx.has
        return dx, dW, dg, (db if has_b else None), (db if has_e else None), None


def dense_wn_train(x, V, g, b, act=NONE, extra_bias=None):
    """dense_wn (cond:437-459), differentiable: (x @ V) * g / ||V|| + b (+ extra_bias)."""
    return _DenseLayerFn.apply(x, V, g, b, extra_bias, act)


def dense_train(x, K, bias=None, act=NONE):
    """plain tf.layers.dense, differentiable."""
    return _DenseLayerFn.apply(x, K, None, bias, None, act)


def _grad_mode(*tensors):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class _CondConvCouplingFn(torch.autograd.Function):
    """Label-conditioned CouplingLayerConv: forward = bruno_coupling_conv_fwd with per-image bias tables and head terms,
    backward = bruno_coupling_conv_bwd_cond (gradients of the variables AND of tab / ls / lt, which autograd carries on
    into the label paths)."""
    @staticmethod
    def forward(ctx, x, ld, V1, g1, Vr, gr, Ks, bs, Kt, bt, tab, ls, lt, layer):
        x, ld, tab, ls, lt = x.contiguous(), ld.contiguous(), tab.contiguous(), ls.contiguous(), lt.contiguous()
        N, H, W, C = x.shape
        R, Fl = layer.num_res_blocks, layer.num_filters
        nl = 2 * R + 1
        wb = _lib.query('bruno_conv_wpack_bytes')
        wfwd = torch.empty(2 * R, wb, dtype=torch.uint8, device=x.device)
        wbwd = torch.empty(2 * R, wb, dtype=torch.uint8, device=x.device)
        _lib.call('bruno_conv_wn_pack', Vr, gr, wfwd, wbwd, 2 * R)
        slabs = layer._saved_slabs(N, H, W, x.device)
        y = torch.empty_like(x)
        ld_out = torch.empty_like(ld)
        _lib.call('bruno_coupling_conv_fwd', x, ld, y, ld_out, N, H, W, C, R, layer.mask_id, 0, V1, g1, tab, wfwd,
                  tab.data_ptr() + 4 * Fl, nl * Fl, Ks, bs, Kt, bt, ls, lt, slabs, int(slabs.shape[0]), 3)
        ctx.layer, ctx.slabs, ctx.wbwd = layer, slabs, wbwd
        layer._generation += 1
        ctx.generation = layer._generation
        ctx.save_for_backward(x, V1, g1, Vr, gr, Ks, bs, Kt, bt, ls)
        return y, ld_out

    @staticmethod
    def backward(ctx, dy, dld):
        x, V1, g1, Vr, gr, Ks, bs, Kt, bt, ls = ctx.saved_tensors
        layer = ctx.layer
        _check_generation(ctx, layer)
        N, H, W, C = x.shape
        R, Fl = layer.num_res_blocks, layer.num_filters
        nl = 2 * R + 1
        dy, dld = dy.contiguous(), dld.contiguous()
        gsl = SLABS.get('grad', nl, N, H, W, x.device)
        dx = torch.empty_like(x)
        dV1, dg1, db1 = torch.empty_like(V1), torch.empty_like(g1), torch.empty_like(g1)
        dVr, dgr, dbr = torch.empty_like(Vr), torch.empty_like(gr), torch.empty_like(gr)
        dKs, dbs, dKt, dbt = torch.empty_like(Ks), torch.empty_like(bs), torch.empty_like(Kt), torch.empty_like(bt)
        dtab = _new_like(x, N, nl * Fl)
        dls, dlt = _new_like(x, N, C), _new_like(x, N, C)
        scratch = _new_like(x, _lib.query('bruno_heads_bwd_scratch_floats', C))
        _lib.call('bruno_coupling_conv_bwd_cond', x, dy, dld, N, H, W, C, R, layer.mask_id, V1, g1, Vr, gr, ctx.wbwd, Ks, bs,
                  Kt, bt, ctx.slabs, gsl, scratch, dx, dV1, dg1, db1, dVr, dgr, dbr, dKs, dbs, dKt, dbt, ls, dtab, nl * Fl,
                  dls, dlt, 3)
        # the conv biases live inside `tab` (tab = label term + conv bias): db1 / dbr reach them through dtab
        return dx, dld, dV1, dg1, dVr, dgr, dKs, dbs, dKt, dbt, dtab, dls, dlt, None


def _ptr_array(tensors):
    """`const float* const*` argument of the C ABI: a host array of device pointers (kept alive by the caller)."""
    import ctypes
    for t in tensors:
        _lib.ptr(t)                                        # CUDA / float32 / contiguous checks
    return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _CondDenseCouplingFn(torch.autograd.Function):
    """Label-conditioned CouplingLayerDense (cond:162-266): forward = bruno_coupling_dense_cond_fwd, backward =
    bruno_coupling_dense_cond_bwd -- hand-written: the gradients of the 20 variables, of x and of the label embedding
    come from one C-ABI call each way (GEMMs on the tcgen05 3xTF32 path, fused scale / bias / activation epilogues)."""
    @staticmethod
    def forward(ctx, x, ld, yl, layer, inverse, *vs):
        xs = x.shape
        x, ld, yl = x.contiguous(), ld.contiguous(), yl.contiguous()
        N, D, U, L = xs[0], int(np.prod(xs[1:])), layer.n_units, yl.shape[1]
        vs = [v.contiguous() for v in vs]
        ws = _new_like(x, _lib.query('bruno_coupling_dense_cond_ws_floats', N, D, U, L))
        y, ld_out = torch.empty_like(x), torch.empty_like(ld)
        _lib.call('bruno_coupling_dense_cond_fwd', x, yl, ld, y, ld_out, N, D, U, L, layer.mask_id, int(inverse), _ptr_array(vs), ws)
        if any(ctx.needs_input_grad):
            if inverse:
                raise RuntimeError('backward through the inverse pass is not implemented (sampling is inference-only)')
            ctx.dims = (N, D, U, L, layer.mask_id)
            ctx.save_for_backward(x, yl, ws, *vs)
        return y, ld_out

    @staticmethod
    def backward(ctx, dy, dld):
        x, yl, ws = ctx.saved_tensors[:3]
        vs = list(ctx.saved_tensors[3:])
        N, D, U, L, mask_id = ctx.dims
        ws2 = _new_like(x, _lib.query('bruno_coupling_dense_cond_bwd_ws_floats', N, D, U, L))
        dx, dyl = torch.empty_like(x), torch.empty_like(yl)
        grads = [torch.empty_like(v) for v in vs]
        _lib.call('bruno_coupling_dense_cond_bwd', x, yl, dy.contiguous(), dld.contiguous(), N, D, U, L, mask_id, _ptr_array(vs), ws,
                  ws2, dx, dyl, _ptr_array(grads))
        return (dx, dld, dyl, None, None) + tuple(grads)


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
        self._slabs = None
        self._generation = 0

    # trainable variables (the label_h bias of a conv is created with use_bias=False -> not trainable, cond:426)
    _TRAINABLE = ('V1', 'g1', 'b1', 'Vr', 'gr', 'br', 'LV', 'Lg', 'Ks', 'bs', 'Kt', 'bt', 'LKs', 'LKt')

    def parameters(self):
        return [] if self.V1 is None else [getattr(self, n) for n in self._TRAINABLE]

    def _saved_slabs(self, N, H, W, device):
        n = 2 * self.num_res_blocks + 1
        key = (N, H, W, str(device))
        if self._slabs is None or self._slabs[0] != key:
            self._slabs = (key, alloc_slabs(n, N, H, W, device))
        return self._slabs[1]

    def _build(self, C, L, device):
        F, R = self.num_filters, self.num_res_blocks
        nl = 2 * R + 1
        # c1 of a conditional coupling is 3x3 in the reference (cond:427 drops filter_size): V1 = HWIO [3,3,C,F] flattened
        self.V1 = _normal(9 * C, F, device=device)
        self.g1, self.b1 = torch.ones(F, device=device), torch.zeros(F, device=device)
        self.Vr = _normal(2 * R, 3, 3, F, F, device=device)
        self.gr, self.br = torch.ones(2 * R, F, device=device), torch.zeros(2 * R, F, device=device)
        # label paths of the 2R+1 weight-normalised convs, concatenated along the output dim: [L, (2R+1)*64]
        self.LV = _normal(L, nl * F, device=device)
        self.Lg, self.Lb = torch.ones(nl * F, device=device), torch.zeros(nl * F, device=device)
        self.Ks, self.bs = torch.zeros(F, C, device=device), torch.zeros(C, device=device)
        self.Kt, self.bt = torch.zeros(F, C, device=device), torch.zeros(C, device=device)
        self.LKs, self.LKt = _normal(L, C, device=device, std=0.3), _normal(L, C, device=device, std=0.3)

    def named_tf_variables(self, prefix='model/', grad=False):
        """(TF variable name, view) pairs; grad=True: the same views of the gradients (None where there is none)."""
        sc = '%s%s/function_s_t_wn/' % (prefix, self.name)
        F, R, C = self.num_filters, self.num_res_blocks, self.V1.shape[0] // 9
        names = ['c1'] + [n for r in range(R) for n in ('c2_%d' % r, 'c3_%d' % r)]

        def T(n):
            t = getattr(self, n)
            return t.grad if grad else t

        def V(t, f):
            return None if t is None else f(t)
        out = []
        for i, nm in enumerate(names):
            out += [(sc + nm + '/label_h/V', V(T('LV'), lambda t: t[:, i * F:(i + 1) * F])),
                    (sc + nm + '/label_h/g', V(T('Lg'), lambda t: t[i * F:(i + 1) * F])),
                    (sc + nm + '/label_h/b', V(T('Lb'), lambda t: t[i * F:(i + 1) * F]))]
            if i == 0:
                out += [(sc + 'c1/conv_h/V', V(T('V1'), lambda t: t.view(3, 3, C, F))), (sc + 'c1/conv_h/g', T('g1')),
                        (sc + 'c1/conv_h/b', T('b1'))]
            else:
                out += [(sc + nm + '/conv_h/V', V(T('Vr'), lambda t: t[i - 1])), (sc + nm + '/conv_h/g', V(T('gr'), lambda t: t[i - 1])),
                        (sc + nm + '/conv_h/b', V(T('br'), lambda t: t[i - 1]))]
        for nm, K, b, LK in (('conv_out_scale', 'Ks', 'bs', 'LKs'), ('conv_out_translation', 'Kt', 'bt', 'LKt')):
            out += [('%s%s/label_h/kernel' % (sc, nm), T(LK)), ('%s%s/%s/kernel' % (sc, nm, nm), V(T(K), lambda t: t.view(1, 1, F, C))),
                    ('%s%s/%s/bias' % (sc, nm, nm), T(b))]
        return out

    def _apply(self, x, ld, y_label, inverse):
        x = x.contiguous()
        N, H, W, C = x.shape
        if self.V1 is None:
            self._build(C, int(y_label.shape[1]), x.device)
        F, R = self.num_filters, self.num_res_blocks
        nl = 2 * R + 1
        conv_b = torch.cat([self.b1.view(1, F), self.br], 0).reshape(-1)
        if _grad_mode(x, ld, y_label, *self.parameters()):
            if inverse:
                raise RuntimeError('backward through the inverse pass is not implemented (sampling is inference-only)')
            tab = dense_wn_train(y_label, self.LV, self.Lg, self.Lb, NONE, extra_bias=conv_b)
            ls = dense_train(y_label, self.LKs)
            lt = dense_train(y_label, self.LKt)
            return _CondConvCouplingFn.apply(x, ld, self.V1, self.g1, self.Vr, self.gr, self.Ks, self.bs, self.Kt, self.bt,
                                             tab, ls, lt, self)
        with torch.no_grad():
            # per-image bias tables [N, (2R+1)*64]: dense_wn(y_label) + b of the conv (cond:422-433)
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
                      self.g1, tab, wfwd, tab.data_ptr() + 4 * F, nl * F, self.Ks, self.bs, self.Kt, self.bt, ls, lt, slabs, 3, 3)
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

    def named_tf_variables(self, prefix='model/', grad=False):
        sc = '%s%s/function_s_t_dense_wn/' % (prefix, self.name)
        return [(sc + k, (v.grad if grad else v)) for k, v in self.P.items()]

    def parameters(self):
        return [] if self.P is None else list(self.P.values())      # every variable of the dense path is trainable

    def _apply(self, x, ld, y_label, inverse):
        xs = x.shape
        x = x.contiguous()
        N, D = xs[0], int(np.prod(xs[1:]))
        if self.P is None:
            self._build(D, int(y_label.shape[1]), x.device)
        vs = list(self.P.values())
        assert len(vs) == 20
        ld_in = ld if ld is not None else torch.zeros(N, device=x.device)
        if inverse and _grad_mode(x, ld, y_label, *vs):
            raise RuntimeError('backward through the inverse pass is not implemented (sampling is inference-only)')
        out, ld_out = _CondDenseCouplingFn.apply(x, ld_in, y_label, self, bool(inverse), *vs)
        return out.reshape(xs), ld_out

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
