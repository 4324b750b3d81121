"""Real NVP coupling stack -- host mirror of the reference's nn_extra_nvp.py.

Same classes, constructor arguments and layer protocol as the reference
(`forward_and_jacobian(x, z, sum_log_det_jacobian)` / `backward(y, z, sum_log_det_jacobian)`, nn_extra_nvp.py:14-19;
LogitLayer / ScaleLayer return 2-tuples, :31,:49), tensors are NHWC torch CUDA tensors.  All arithmetic runs in
libbruno_sm100.so (hand-written sm_100a kernels) through the C ABI; every layer is a torch.autograd.Function whose
backward is a hand-written CUDA backward, not autograd of torch ops.  There is no CPU / eager fallback.
"""
import math

import numpy as np
import torch

from . import _lib
from .slab import alloc_slabs

MASKS = {'checkerboard0': 0, 'checkerboard1': 1, 'channel0': 2, 'channel1': 3, 'even': 4, 'odd': 5}

elu = torch.nn.functional.elu          # the only nonlinearity the target configs use (bn2:17)

_INIT_MODE = [False]                   # arg_scope([conv2d_wn, dense_wn], init=init) of the reference (bn2:62)

# Precision of the 3x3 convolutions of CouplingLayerConv:
#   'bf16' (default, production): tcgen05 tensor cores, bf16 operands / activations, fp32 accumulation.
#   'fp32' (verification, inference only): the same layer evaluated entirely in fp32 on CUDA cores (nine tap GEMMs per
#          convolution through bruno_sgemm on fp32 slabs) -- slow, used by the parity tests to show that the only
#          difference between the production path and the reference arithmetic is the declared reduced precision.
#          The mode also routes every dense-path GEMM to the strict-fp32 FFMA kernel (bruno_debug_gemm_variant(0); the
#          production GEMM is the 3xTF32 tensor-core split, ~8x the rounding error of an fp32 FMA chain, which the
#          ill-conditioned Gaussian latent log-density of a sampled image amplifies to 6e-4 .. 1.2e-3 relative against
#          5e-6 for the strict path, tests/test_reference_fixtures_gpu.py::test_sampling_branch_matches_reference).
_PRECISION = ['bf16']


class precision(object):
    """with nn_extra_nvp.precision('fp32'): ..."""

    def __init__(self, p):
        assert p in ('bf16', 'fp32')
        self.p = p
        self.prev_variant = None

    def __enter__(self):
        self.prev = _PRECISION[0]
        _PRECISION[0] = self.p
        lib = _lib.load()
        if self.p == 'fp32' and hasattr(lib, 'bruno_debug_gemm_variant'):    # absent in a -DBRUNO_NO_DEBUG build
            self.prev_variant = lib.bruno_debug_gemm_variant(-1)
            lib.bruno_debug_gemm_variant(0)

    def __exit__(self, *a):
        _PRECISION[0] = self.prev
        if self.prev_variant is not None:
            _lib.load().bruno_debug_gemm_variant(self.prev_variant)
            self.prev_variant = None


class init_scope(object):
    """with init_scope(init): ...  -- the reference's arg_scope(init=init)."""

    def __init__(self, init):
        self.init = bool(init)

    def __enter__(self):
        self.prev = _INIT_MODE[0]
        _INIT_MODE[0] = self.init

    def __exit__(self, *a):
        _INIT_MODE[0] = self.prev


def int_shape(x):                                                           # nvp:10-11
    return list(map(int, x.shape))


def _new_like(x, *shape):
    return torch.empty(*shape, dtype=torch.float32, device=x.device)


class _SlabPool(object):
    """Zero-initialised activation slabs are allocated once per geometry and reused (kernels keep padding zero)."""

    def __init__(self):
        self.pool = {}

    def get(self, tag, n, N, H, W, device):
        key = (tag, n, N, H, W, str(device))
        if key not in self.pool:
            self.pool[key] = alloc_slabs(n, N, H, W, device)
        return self.pool[key]

    def clear(self):
        self.pool.clear()


SLABS = _SlabPool()

# Direct gradient sinks.  After BrunoModel.flatten_parameters every variable owns a preset .grad view into the flat
# gradient buffer; letting autograd accumulate into it costs one elementwise add launch per variable per step (166 for
# bn2_omniglot).  With direct sinks on, the backward kernels WRITE the variable gradients straight into those views and
# hand autograd None.  Valid only under the trainer's contract: the buffer is zeroed every step and every variable is
# used once per backward (no micro-batch accumulation) -- the trainer switches it on, everything else leaves it off.
_DIRECT_GRADS = [False]


def direct_grads(on):
    _DIRECT_GRADS[0] = bool(on)


# Called by the coupling backward once the kernels that write a layer's variable gradients have been enqueued (the trainer
# uses it to start the NCCL allreduce of that layer's slice of the flat gradient buffer while the backward goes on).
_GRADS_READY_HOOK = [None]


def on_layer_grads_ready(fn):
    _GRADS_READY_HOOK[0] = fn


def _notify_grads_ready(layer):
    if _GRADS_READY_HOOK[0] is not None:
        _GRADS_READY_HOOK[0](layer)


def _grad_sinks(layer, names):
    """-> (buffers the kernel writes, values returned to autograd) for the variables `names` of `layer`."""
    bufs, rets = [], []
    for n in names:
        p = getattr(layer, n)
        if _DIRECT_GRADS[0] and p.grad is not None and p.grad.is_contiguous() and p.grad.shape == p.shape:
            bufs.append(p.grad)
            rets.append(None)
        else:
            t = torch.empty_like(p)
            bufs.append(t)
            rets.append(t)
    return bufs, rets


def _check_generation(ctx, layer):
    """The activations a conv coupling saves for its backward are kept in one buffer per layer (6.3 GB per model at
    B=32, T=20: not duplicated per call).  A second grad-enabled forward through the same layer before the backward of
    the first (gradient accumulation over micro-batches, two losses, autograd.grad on an older graph) has overwritten
    them -- fail loudly instead of returning gradients of the wrong activations."""
    if ctx.generation != layer._generation:
        raise RuntimeError('%s: backward of a stale graph -- the layer ran %d grad-enabled forward(s) after the one being '
                           'differentiated and its saved activations were overwritten; call backward before the next '
                           'forward' % (layer.name, layer._generation - ctx.generation))


class Layer(object):                                                        # nvp:14-19
    def forward_and_jacobian(self, x, z, sum_log_det_jacobian):
        raise NotImplementedError(str(type(self)))

    def backward(self, y, z, sum_log_det_jacobian):
        raise NotImplementedError(str(type(self)))

    def parameters(self):
        return []

    def named_tf_variables(self, prefix='model/'):
        return []


def _pre(op, x, ld, scale, alpha):
    xs = int_shape(x)
    x = x.contiguous()
    N, D = xs[0], int(np.prod(xs[1:]))
    y = torch.empty_like(x)
    ld_out = _new_like(x, N)
    _lib.call('bruno_pre', op, x, None if ld is None else ld.contiguous(), y, ld_out, N, D, float(scale), float(alpha), 0)
    return y, ld_out


class LogitLayer(Layer):                                                    # nvp:22-38
    def __init__(self, alpha=1e-5):
        self.alpha = alpha

    def forward_and_jacobian(self, x, z, sum_log_det_jacobian):
        return _pre(1, x.detach(), sum_log_det_jacobian, 256., self.alpha)

    def backward(self, y, z, sum_log_det_jacobian):
        return _pre(2, y.detach(), sum_log_det_jacobian, 256., self.alpha)


class ScaleLayer(Layer):                                                    # nvp:41-55
    def __init__(self, scale=256.):
        self.scale = scale

    def forward_and_jacobian(self, x, z, sum_log_det_jacobian):
        return _pre(0, x.detach(), sum_log_det_jacobian, self.scale, 1e-5)

    def backward(self, y, z, sum_log_det_jacobian):
        return _pre(3, y.detach(), sum_log_det_jacobian, self.scale, 1e-5)


def preprocess_forward(x, scale=256., alpha=1e-5):
    """ScaleLayer then LogitLayer fused in one kernel (bn2:82-83): returns (y, log_det) with log_det starting at 0."""
    return _pre(4, x.detach(), None, scale, alpha)


def preprocess_backward(y, ld, scale=256., alpha=1e-5):
    """LogitLayer.backward then ScaleLayer.backward fused (bn2:148-149)."""
    return _pre(5, y.detach(), ld, scale, alpha)


# ------------------------------------------------------------------------------------------------------------------
# conv coupling
# ------------------------------------------------------------------------------------------------------------------
class _ConvCouplingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ld, V1, g1, b1, Vr, gr, br, Ks, bs, Kt, bt, layer, inverse, track):
        x = x.contiguous()
        ld = ld.contiguous()
        N, H, W, C = x.shape
        R = layer.num_res_blocks
        need_grad = track and any(ctx.needs_input_grad)
        wb = _lib.query('bruno_conv_wpack_bytes')
        wfwd = torch.empty(2 * R, wb, dtype=torch.uint8, device=x.device)
        wbwd = torch.empty(2 * R, wb, dtype=torch.uint8, device=x.device) if need_grad else None
        _lib.call('bruno_conv_wn_pack', Vr, gr, wfwd, wbwd, 2 * R)
        if need_grad:
            slabs = layer._saved_slabs(N, H, W, x.device)
        else:
            slabs = SLABS.get('pingpong', 3, N, H, W, x.device)
        y = torch.empty_like(x)
        ld_out = torch.empty_like(ld)
        _lib.call('bruno_coupling_conv_fwd', x, ld, y, ld_out, N, H, W, C, R, layer.mask_id, int(inverse), V1, g1, b1, wfwd,
                  br, 0, Ks, bs, Kt, bt, None, None, slabs, int(slabs.shape[0]), 1)
        if need_grad:
            if inverse:
                raise RuntimeError('backward through the inverse pass is not implemented (sampling is inference-only)')
            ctx.layer = layer
            ctx.slabs = slabs
            ctx.wbwd = wbwd
            # the saved activations live in ONE per-layer buffer: a later grad-enabled forward overwrites them
            layer._generation += 1
            ctx.generation = layer._generation
            ctx.save_for_backward(x, V1, g1, Vr, gr, Ks, bs, Kt, bt)
        return y, ld_out

    @staticmethod
    def backward(ctx, dy, dld):
        x, V1, g1, Vr, gr, Ks, bs, Kt, bt = ctx.saved_tensors
        layer = ctx.layer
        _check_generation(ctx, layer)
        N, H, W, C = x.shape
        R = layer.num_res_blocks
        dy = dy.contiguous()
        dld = dld.contiguous()
        gsl = SLABS.get('grad', 2 * R + 1, N, H, W, x.device)
        dx = torch.empty_like(x)
        (dV1, dg1, db1, dVr, dgr, dbr, dKs, dbs, dKt, dbt), rets = _grad_sinks(layer, layer._PNAMES)
        scratch = _new_like(x, _lib.query('bruno_heads_bwd_scratch_floats', C))
        _lib.call('bruno_coupling_conv_bwd', x, dy, dld, N, H, W, C, R, layer.mask_id, V1, g1, Vr, gr, ctx.wbwd, Ks, bs, Kt, bt,
                  ctx.slabs, gsl, scratch, dx, dV1, dg1, db1, dVr, dgr, dbr, dKs, dbs, dKt, dbt)
        _notify_grads_ready(layer)
        return (dx, dld) + tuple(rets) + (None, None, None)


class CouplingLayerConv(Layer):                                             # nvp:58-187
    def __init__(self, mask_type, name='CouplingLayer', nonlinearity=elu, weight_norm=True, num_filters=64,
                 num_res_blocks=8):
        if mask_type not in ('checkerboard0', 'checkerboard1', 'channel0', 'channel1'):
            raise AssertionError('unknown mask type %s' % mask_type)        # nvp:145
        if not weight_norm:
            raise NotImplementedError('the B200 path implements the weight-normalised s/t network (function_s_t_wn); '
                                      'the non-WN variant is only used by config_failed/')
        if num_filters != 64:
            raise NotImplementedError('tcgen05 conv kernels are built for num_filters = 64 (got %d)' % num_filters)
        if nonlinearity is not elu:
            raise NotImplementedError('only ELU is fused into the conv epilogues (all target configs use it)')
        self.mask_type = mask_type
        self.mask_id = MASKS[mask_type]
        self.name = name
        self.nonlinearity = nonlinearity
        self.weight_norm = weight_norm
        self.num_filters = num_filters
        self.num_res_blocks = num_res_blocks
        self.V1 = None
        self._slabs = None
        self._generation = 0

    # ---- variables (created on first use, like tf.get_variable under make_template) --------------------
    def _build(self, C, device):
        F, R = self.num_filters, self.num_res_blocks
        def normal(*s):                                                      # nvp:382 (stddev 0.05)
            return (torch.randn(*s, device=device) * 0.05).requires_grad_(True)
        self.V1 = normal(C, F)
        self.g1 = torch.ones(F, device=device, requires_grad=True)
        self.b1 = torch.zeros(F, device=device, requires_grad=True)
        self.Vr = normal(2 * R, 3, 3, F, F)
        self.gr = torch.ones(2 * R, F, device=device, requires_grad=True)
        self.br = torch.zeros(2 * R, F, device=device, requires_grad=True)
        self.Ks = torch.zeros(F, C, device=device, requires_grad=True)      # zero-initialised heads, nvp:129-138
        self.bs = torch.zeros(C, device=device, requires_grad=True)
        self.Kt = torch.zeros(F, C, device=device, requires_grad=True)
        self.bt = torch.zeros(C, device=device, requires_grad=True)

    _PNAMES = ('V1', 'g1', 'b1', 'Vr', 'gr', 'br', 'Ks', 'bs', 'Kt', 'bt')

    def parameters(self):
        return [] if self.V1 is None else [getattr(self, n) for n in self._PNAMES]

    def named_parameters(self):
        return [] if self.V1 is None else [(self.name + '/' + n, getattr(self, n)) for n in self._PNAMES]

    def _set_parameters(self, tensors):
        for n, t in zip(self._PNAMES, tensors):
            setattr(self, n, t)

    def named_tf_variables(self, prefix='model/'):
        """[(TF variable name, tensor view with the TF shape)] for checkpoint interchange."""
        sc = '%s%s/function_s_t_wn/' % (prefix, self.name)
        C = self.V1.shape[0]
        out = [(sc + 'c1/V', self.V1.view(1, 1, C, -1)), (sc + 'c1/g', self.g1), (sc + 'c1/b', self.b1)]
        for r in range(self.num_res_blocks):
            for j, nm in enumerate(('c2_%d' % r, 'c3_%d' % r)):
                out += [(sc + nm + '/V', self.Vr[2 * r + j]), (sc + nm + '/g', self.gr[2 * r + j]),
                        (sc + nm + '/b', self.br[2 * r + j])]
        out += [(sc + 'conv_out_scale/kernel', self.Ks.view(1, 1, -1, C)), (sc + 'conv_out_scale/bias', self.bs),
                (sc + 'conv_out_translation/kernel', self.Kt.view(1, 1, -1, C)), (sc + 'conv_out_translation/bias', self.bt)]
        return out

    def _saved_slabs(self, N, H, W, device):
        n = 2 * self.num_res_blocks + 1
        key = (N, H, W, str(device))
        if self._slabs is None or self._slabs[0] != key:
            self._slabs = (key, alloc_slabs(n, N, H, W, device))
        return self._slabs[1]

    def _apply_fp32(self, x, ld, inverse):
        """fp32 verification path (see _PRECISION): c1 -> 2R convolutions as 9 tap GEMMs each -> heads, all fp32.
        Under init_scope(True) it also performs the data-dependent init (nvp:394-398) from the fp32 pre-activations, so
        that the init can be compared with the reference's at fp32 accuracy (the production init uses the moments of the
        bf16 path's own activations, bruno_b200/ddinit.py)."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in (x, ld)):
            raise RuntimeError("precision('fp32') is an inference-only verification mode")
        from .slab import SlabGeom, _pixel_rows
        init = _INIT_MODE[0]
        with torch.no_grad():
            x, ld = x.contiguous(), ld.contiguous()
            N, H, W, C = x.shape
            g = SlabGeom(N, H, W)
            key = ('f32', N, H, W, str(x.device))
            if key not in SLABS.pool:
                SLABS.pool[key] = torch.zeros(3, g.rows, 64, device=x.device)
            buf = SLABS.pool[key]
            M = g.NT * 128
            _lib.call('bruno_c1_fwd', x, self.V1, self.g1, self.b1, 0, buf[0], 1, N, H, W, C, self.mask_id, 1)
            moments = []
            rows = _pixel_rows(g, x.device) if init else None

            def conv(src, dst, l, skip, act):
                V = self.Vr[l].detach().contiguous()
                scale = (self.gr[l] / torch.sqrt(torch.clamp((V * V).sum(dim=(0, 1, 2)), min=1e-12))).contiguous()
                for tap in range(9):
                    shift = (tap // 3 - 1) * g.P + (tap % 3 - 1)
                    _lib.call('bruno_sgemm', 0, 0, M, 64, 64, src.data_ptr() + 4 * 64 * (g.HALO + shift), 64,
                              V.data_ptr() + 4 * 64 * 64 * tap, 64, dst.data_ptr() + 4 * 64 * g.HALO, 64, scale, None, 0,
                              1 if tap > 0 else 0)
                if init:                                   # moments of the pre-activation conv + b over (N, H, W)
                    pre = dst[rows].double() + self.br[l].double()
                    moments.append((pre.mean(0), pre.var(0, unbiased=False)))
                _lib.call('bruno_conv_epilogue_f32', dst, self.br[l].detach().contiguous(), skip, int(act), N, H, W)

            hi = 0
            for r in range(self.num_res_blocks):
                ui, ni = (hi + 1) % 3, (hi + 2) % 3
                conv(buf[hi], buf[ui], 2 * r, None, True)
                conv(buf[ui], buf[ni], 2 * r + 1, buf[hi], True)
                hi = ni
            y = torch.empty_like(x)
            ld_out = torch.empty_like(ld)
            _lib.call('bruno_heads_coupling', x, buf[hi], self.Ks, self.bs, self.Kt, self.bt, None, None, ld, y, ld_out, 1,
                      N, H, W, C, self.mask_id, int(inverse))
            if init:
                from .ddinit import conv_mask, assign_from_moments
                b = conv_mask(self.mask_type, H, W, C, x.device)
                nrm = torch.sqrt(torch.clamp((self.V1.double() ** 2).sum(0), min=1e-12))
                pre1 = (x * b).reshape(-1, C).double() @ (self.V1.double() * (self.g1.double() / nrm)) + self.b1.double()
                assign_from_moments(self.g1, self.b1, pre1.mean(0), pre1.var(0, unbiased=False))
                for l, (m, v) in enumerate(moments):
                    assign_from_moments(self.gr[l], self.br[l], m, v)
        return y, ld_out

    def _apply(self, x, ld, inverse):
        if self.V1 is None:
            self._build(int(x.shape[3]), x.device)
        if _PRECISION[0] == 'fp32':
            return self._apply_fp32(x, ld, inverse)          # handles init_scope itself
        out = _ConvCouplingFn.apply(x, ld, self.V1, self.g1, self.b1, self.Vr, self.gr, self.br, self.Ks, self.bs,
                                    self.Kt, self.bt, self, inverse, torch.is_grad_enabled())
        if _INIT_MODE[0]:                      # outputs above use the OLD g, b (nvp:396-398: tf.identity(x))
            from .ddinit import conv_coupling_init
            conv_coupling_init(self, x)
        return out

    def forward_and_jacobian(self, x, z, sum_log_det_jacobian):            # nvp:162-174
        y, ld = self._apply(x, sum_log_det_jacobian, False)
        return y, z, ld

    def backward(self, y, z, sum_log_det_jacobian):                        # nvp:176-187
        x, ld = self._apply(y, sum_log_det_jacobian, True)
        return x, z, ld


# ------------------------------------------------------------------------------------------------------------------
# dense coupling
# ------------------------------------------------------------------------------------------------------------------
class _DenseCouplingFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, ld, V1, g1, b1, V2, g2, b2, Ks, bs, Kt, bt, layer, inverse, track):
        xs = x.shape
        x = x.contiguous()
        ld = ld.contiguous()
        N, D, U = xs[0], int(np.prod(xs[1:])), V1.shape[1]
        ws = _new_like(x, _lib.query('bruno_coupling_dense_ws_floats', N, D, U))
        y = torch.empty_like(x)
        ld_out = torch.empty_like(ld)
        _lib.call('bruno_coupling_dense_fwd', x, ld, y, ld_out, N, D, U, layer.mask_id, int(inverse), V1, g1, b1, V2, g2, b2,
                  Ks, bs, Kt, bt, ws)
        need_grad = track and any(ctx.needs_input_grad)
        if need_grad:
            if inverse:
                raise RuntimeError('backward through the inverse pass is not implemented (sampling is inference-only)')
            ctx.layer = layer
            ctx.save_for_backward(x, V1, g1, b1, V2, g2, b2, Ks, Kt, ws)
        return y, ld_out

    @staticmethod
    def backward(ctx, dy, dld):
        x, V1, g1, b1, V2, g2, b2, Ks, Kt, ws = ctx.saved_tensors
        xs = x.shape
        N, D, U = xs[0], int(np.prod(xs[1:])), V1.shape[1]
        ws2 = _new_like(x, _lib.query('bruno_coupling_dense_bwd_ws_floats', N, D, U))
        dx = torch.empty_like(x)
        (dV1, dg1, db1, dV2, dg2, db2, dKs, dbs, dKt, dbt), rets = _grad_sinks(ctx.layer, ctx.layer._PNAMES)
        _lib.call('bruno_coupling_dense_bwd', x, dy.contiguous(), dld.contiguous(), N, D, U, ctx.layer.mask_id, V1, g1, b1,
                  V2, g2, b2, Ks, Kt, ws, ws2, dx, dV1, dg1, db1, dV2, dg2, db2, dKs, dbs, dKt, dbt)
        _notify_grads_ready(ctx.layer)
        return (dx, dld) + tuple(rets) + (None, None, None)


class CouplingLayerDense(Layer):                                            # nvp:190-274
    def __init__(self, mask_type, name='CouplingDense', nonlinearity=elu, n_units=1024, weight_norm=True):
        if mask_type not in ('even', 'odd'):
            raise AssertionError('unknown mask type %s' % mask_type)        # nvp:201
        if not weight_norm:
            raise NotImplementedError('only the weight-normalised dense s/t network is implemented')
        if nonlinearity is not elu:
            raise NotImplementedError('only ELU is fused into the dense epilogues')
        self.mask_type = mask_type
        self.mask_id = MASKS[mask_type]
        self.name = name
        self.nonlinearity = nonlinearity
        self.weight_norm = weight_norm
        self.n_units = n_units
        self.V1 = None

    def _build(self, D, device):
        U = self.n_units
        def normal(*s):                                                      # nvp:411 (stddev 0.05)
            return (torch.randn(*s, device=device) * 0.05).requires_grad_(True)
        self.V1, self.V2 = normal(D, U), normal(U, U)
        self.g1 = torch.ones(U, device=device, requires_grad=True)
        self.g2 = torch.ones(U, device=device, requires_grad=True)
        self.b1 = torch.zeros(U, device=device, requires_grad=True)
        self.b2 = torch.zeros(U, device=device, requires_grad=True)
        self.Ks = torch.zeros(U, D, device=device, requires_grad=True)      # zero-initialised, nvp:262-270
        self.bs = torch.zeros(D, device=device, requires_grad=True)
        self.Kt = torch.zeros(U, D, device=device, requires_grad=True)
        self.bt = torch.zeros(D, device=device, requires_grad=True)

    _PNAMES = ('V1', 'g1', 'b1', 'V2', 'g2', 'b2', 'Ks', 'bs', 'Kt', 'bt')

    def parameters(self):
        return [] if self.V1 is None else [getattr(self, n) for n in self._PNAMES]

    def named_parameters(self):
        return [] if self.V1 is None else [(self.name + '/' + n, getattr(self, n)) for n in self._PNAMES]

    def _set_parameters(self, tensors):
        for n, t in zip(self._PNAMES, tensors):
            setattr(self, n, t)

    def named_tf_variables(self, prefix='model/'):
        sc = '%s%s/function_s_t_dense_wn/' % (prefix, self.name)
        return [(sc + 'd1/V', self.V1), (sc + 'd1/g', self.g1), (sc + 'd1/b', self.b1),
                (sc + 'd2/V', self.V2), (sc + 'd2/g', self.g2), (sc + 'd2/b', self.b2),
                (sc + 'd_scale/kernel', self.Ks), (sc + 'd_scale/bias', self.bs),
                (sc + 'd_translate/kernel', self.Kt), (sc + 'd_translate/bias', self.bt)]

    def _apply(self, x, ld, inverse):
        if self.V1 is None:
            self._build(int(np.prod(x.shape[1:])), x.device)
        out = _DenseCouplingFn.apply(x, ld, self.V1, self.g1, self.b1, self.V2, self.g2, self.b2, self.Ks, self.bs,
                                     self.Kt, self.bt, self, inverse, torch.is_grad_enabled())
        if _INIT_MODE[0]:                      # outputs above use the OLD g, b (nvp:424-426)
            from .ddinit import dense_coupling_init
            dense_coupling_init(self, x)
        return out

    def forward_and_jacobian(self, x, z, sum_log_det_jacobian):
        y, ld = self._apply(x, sum_log_det_jacobian, False)
        return y, z, ld

    def backward(self, y, z, sum_log_det_jacobian):
        x, ld = self._apply(y, sum_log_det_jacobian, True)
        return x, z, ld


# ------------------------------------------------------------------------------------------------------------------
# squeeze / factor-out
# ------------------------------------------------------------------------------------------------------------------
def _s2d(x, inverse):
    x = x.contiguous()
    N, H, W, C = x.shape
    if inverse:
        out = _new_like(x, N, H * 2, W * 2, C // 4)
        _lib.call('bruno_space_to_depth', x, out, N, H * 2, W * 2, C // 4, 1)
    else:
        out = _new_like(x, N, H // 2, W // 2, C * 4)
        _lib.call('bruno_space_to_depth', x, out, N, H, W, C, 0)
    return out


class _SpaceToDepthFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, inverse):
        ctx.inverse = inverse
        return _s2d(x, inverse)

    @staticmethod
    def backward(ctx, dy):
        return _s2d(dy, not ctx.inverse), None


def _chan_slice(x, off, n):
    x = x.contiguous()
    C = x.shape[-1]
    out = _new_like(x, *(list(x.shape[:-1]) + [n]))
    _lib.call('bruno_channel_copy', x, out, x.numel() // C, C, off, n, 0, n)
    return out


class _ChannelSliceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, off, n):
        ctx.off, ctx.C = off, x.shape[-1]
        return _chan_slice(x, off, n)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        n = dy.shape[-1]
        dx = torch.zeros(*(list(dy.shape[:-1]) + [ctx.C]), dtype=torch.float32, device=dy.device)
        _lib.call('bruno_channel_copy', dy, dx, dy.numel() // n, n, 0, ctx.C, ctx.off, n)
        return dx, None, None


def _chan_concat(a, b):
    a, b = a.contiguous(), b.contiguous()
    ca, cb = a.shape[-1], b.shape[-1]
    out = _new_like(a, *(list(a.shape[:-1]) + [ca + cb]))
    rows = a.numel() // ca
    _lib.call('bruno_channel_copy', a, out, rows, ca, 0, ca + cb, 0, ca)
    _lib.call('bruno_channel_copy', b, out, rows, cb, 0, ca + cb, ca, cb)
    return out


class _ChannelConcatFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        ctx.ca, ctx.cb = a.shape[-1], b.shape[-1]
        return _chan_concat(a, b)

    @staticmethod
    def backward(ctx, dy):
        return _chan_slice(dy, 0, ctx.ca), _chan_slice(dy, ctx.ca, ctx.cb)


def channel_slice(x, off, n):
    return _ChannelSliceFn.apply(x, off, n)


def channel_concat(a, b):
    return _ChannelConcatFn.apply(a, b)


class SqueezingLayer(Layer):                                                # nvp:277-298
    def __init__(self, name="Squeeze"):
        self.name = name

    def forward_and_jacobian(self, x, z, sum_log_det_jacobian):
        xs = int_shape(x)
        assert xs[1] % 2 == 0 and xs[2] % 2 == 0
        y = _SpaceToDepthFn.apply(x, False)
        if z is not None:
            z = _SpaceToDepthFn.apply(z, False)
        return y, z, sum_log_det_jacobian

    def backward(self, y, z, sum_log_det_jacobian):
        ys = int_shape(y)
        assert ys[3] % 4 == 0
        x = _SpaceToDepthFn.apply(y, True)
        if z is not None:
            z = _SpaceToDepthFn.apply(z, True)
        return x, z, sum_log_det_jacobian


class FactorOutLayer(Layer):                                                # nvp:301-346
    def __init__(self, scale, name='FactorOut'):
        self.scale = scale
        self.name = name

    def forward_and_jacobian(self, x, z, sum_log_det_jacobian):
        xs = int_shape(x)
        split = xs[3] // 2
        new_z = channel_slice(x, 0, split)
        x = channel_slice(x, split, xs[3] - split)
        z = channel_concat(z, new_z) if z is not None else new_z
        return x, z, sum_log_det_jacobian

    def backward(self, y, z, sum_log_det_jacobian):
        zs = int_shape(z)
        if y is None:
            split = zs[3] // (2 ** self.scale)
        else:
            split = int_shape(y)[3]
        new_y = channel_slice(z, zs[3] - split, split)
        z = channel_slice(z, 0, zs[3] - split) if zs[3] - split > 0 else None
        assert int_shape(new_y)[3] == split                                 # nvp:339
        x = channel_concat(new_y, y) if y is not None else new_y
        return x, z, sum_log_det_jacobian
