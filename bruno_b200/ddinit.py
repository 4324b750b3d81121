"""Data-dependent initialisation of the weight-normalised layers (nn_extra_nvp.py:394-398, 422-426; run once at
iteration 0, config_rnn/train.py:178-182).  One-off host logic: the pre-activations come from the CUDA kernels (mode 4
of bruno_conv3x3 / the fp32 workspace of the dense coupling), the per-channel moments are plain torch reductions.
Like the reference, the layer's OUTPUT during the init pass is computed with the OLD g, b; they are re-assigned after."""
import torch

from . import _lib
from .slab import slab_to_nhwc, alloc_slabs


def _assign(g, b, pre, dims):
    m = pre.mean(dim=dims)
    v = pre.var(dim=dims, unbiased=False)
    scale_init = 1.0 / torch.sqrt(v + 1e-10)
    g.mul_(scale_init)                                   # g.assign(g * scale_init)
    b.add_(-m * scale_init)                              # b.assign_add(-m_init * scale_init)


def assign_from_moments(g, b, m, v):
    """g.assign(g * scale_init); b.assign_add(-m_init * scale_init) with scale_init = 1 / sqrt(v_init + 1e-10) (nvp:396-397)."""
    scale_init = (1.0 / torch.sqrt(v + 1e-10)).to(g.dtype)
    g.mul_(scale_init)
    b.add_(-(m.to(g.dtype)) * scale_init)


def conv_mask(mt, H, W, C, device):
    """b of CouplingLayerConv.get_mask (nvp:143-160) as an [H, W, C] float tensor."""
    return _conv_mask(mt, H, W, C, device)


def conv_coupling_init(layer, x):
    with torch.no_grad():
        x = x.contiguous()
        N, H, W, C = x.shape
        R = layer.num_res_blocks
        slabs = layer._saved_slabs(N, H, W, x.device)
        wb = _lib.query('bruno_conv_wpack_bytes')
        wfwd = torch.empty(2 * R, wb, dtype=torch.uint8, device=x.device)
        _lib.call('bruno_conv_wn_pack', layer.Vr, layer.gr, wfwd, None, 2 * R)
        ld = torch.zeros(N, device=x.device)
        y = torch.empty_like(x)
        _lib.call('bruno_coupling_conv_fwd', x, ld, y, torch.empty_like(ld), N, H, W, C, R, layer.mask_id, 0, layer.V1,
                  layer.g1, layer.b1, wfwd, layer.br, 0, layer.Ks, layer.bs, layer.Kt, layer.bt, None, None, slabs, int(slabs.shape[0]), 1)
        # c1: pre-activation on the masked input
        hh = torch.arange(H, device=x.device).view(H, 1, 1)
        ww = torch.arange(W, device=x.device).view(1, W, 1)
        cc = torch.arange(C, device=x.device).view(1, 1, C)
        mt = layer.mask_type
        if mt.startswith('checkerboard'):
            b = ((hh + ww) % 2).float().expand(H, W, C)
            b = 1 - b if mt == 'checkerboard1' else b
        else:
            b = ((cc < C // 2) if mt == 'channel0' else (cc >= C // 2)).float().expand(H, W, C)
        nrm = torch.sqrt(torch.clamp((layer.V1 ** 2).sum(0), min=1e-12))
        W1 = layer.V1 * (layer.g1 / nrm)
        pre = (x * b).reshape(-1, C) @ W1 + layer.b1
        new = [(layer.g1, layer.b1, pre, (0,))]
        tmp = alloc_slabs(1, N, H, W, x.device)[0]
        pres = []
        for l in range(2 * R):
            _lib.call('bruno_conv3x3', 4, slabs[l], wfwd[l], layer.br[l].contiguous(), 0, None, None, tmp, N, H, W)
            pres.append(slab_to_nhwc(tmp, N, H, W))
        _assign(layer.g1, layer.b1, pre, (0,))
        for l in range(2 * R):
            _assign(layer.gr[l], layer.br[l], pres[l], (0, 1, 2))


def _dense_ws_offsets(N, D, U):
    def r4(n):
        return (n + 3) // 4 * 4
    o, out = 0, {}
    for name, n in (('sc1', U), ('sc2', U), ('in1', U), ('in2', U), ('xm', N * D), ('pre1', N * U), ('pre2', N * U), ('h1', N * U), ('h2', N * U),
                    ('s_pre', N * D), ('t_pre', N * D)):
        out[name] = (o, n)
        o += r4(n)
    return out


def dense_coupling_init(layer, x):
    with torch.no_grad():
        xs = x.shape
        x = x.contiguous()
        N, U = xs[0], layer.n_units
        D = x.numel() // N
        ws = torch.empty(_lib.query('bruno_coupling_dense_ws_floats', N, D, U), device=x.device)
        ld = torch.zeros(N, device=x.device)

        def run():
            _lib.call('bruno_coupling_dense_fwd', x, ld, torch.empty_like(x), torch.empty_like(ld), N, D, U, layer.mask_id, 0,
                      layer.V1, layer.g1, layer.b1, layer.V2, layer.g2, layer.b2, layer.Ks, layer.bs, layer.Kt, layer.bt, ws)
        off = _dense_ws_offsets(N, D, U)
        run()
        o, n = off['pre1']
        pre1 = ws[o:o + n].view(N, U).clone()
        o, n = off['pre2']
        pre2 = ws[o:o + n].view(N, U).clone()
        _assign(layer.g1, layer.b1, pre1, (0,))
        _assign(layer.g2, layer.b2, pre2, (0,))


# ------------------------------------------------------------------------------------------------------------------
# label-conditioned layers (nn_extra_nvp_conditional.py:396-471): every weight-normalised conv is conv_h (WN conv, own
# g, b) + label_h (dense_wn of the label embedding, own g, b); each of the two is initialised from the moments of ITS
# OWN pre-activation.  The layer output of the init pass is computed with the old values (like the reference).
# ------------------------------------------------------------------------------------------------------------------
def _conv_mask(mt, H, W, C, device):
    hh = torch.arange(H, device=device).view(H, 1, 1)
    ww = torch.arange(W, device=device).view(1, W, 1)
    cc = torch.arange(C, device=device).view(1, 1, C)
    if mt.startswith('checkerboard'):
        b = ((hh + ww) % 2).float().expand(H, W, C)
        return 1 - b if mt == 'checkerboard1' else b
    return ((cc < C // 2) if mt == 'channel0' else (cc >= C // 2)).float().expand(H, W, C)


def cond_conv_coupling_init(layer, x, ld, y_label):
    from .nn_extra_nvp_conditional import _sgemm
    import torch.nn.functional as Fn
    with torch.no_grad():
        x = x.contiguous()
        N, H, W, C = x.shape
        if layer.V1 is None:
            layer._build(C, int(y_label.shape[1]), x.device)
        F, R = layer.num_filters, layer.num_res_blocks
        nl = 2 * R + 1
        # label paths: pre-activations of the 2R+1 dense_wn(y_label) with the old Lg, Lb
        h_pre = _sgemm(y_label, layer.LV, nl * F) * (layer.Lg / torch.sqrt((layer.LV ** 2).sum(0))) + layer.Lb
        conv_b = torch.cat([layer.b1.view(1, F), layer.br], 0).reshape(-1)
        tab = (h_pre + conv_b).contiguous()
        ls = _sgemm(y_label, layer.LKs, C)
        lt = _sgemm(y_label, layer.LKt, C)
        slabs = layer._saved_slabs(N, H, W, x.device)
        wb = _lib.query('bruno_conv_wpack_bytes')
        wfwd = torch.empty(2 * R, wb, dtype=torch.uint8, device=x.device)
        _lib.call('bruno_conv_wn_pack', layer.Vr, layer.gr, wfwd, None, 2 * R)
        y = torch.empty_like(x)
        ld_out = torch.empty_like(ld)
        _lib.call('bruno_coupling_conv_fwd', x, ld.contiguous(), y, ld_out, N, H, W, C, R, layer.mask_id, 0, layer.V1, layer.g1,
                  tab, wfwd, tab.data_ptr() + 4 * F, nl * F, layer.Ks, layer.bs, layer.Kt, layer.bt, ls, lt, slabs,
                  int(slabs.shape[0]), 3)
        # conv_h pre-activations: the convolution with the conv's own bias only
        b = _conv_mask(layer.mask_type, H, W, C, x.device)
        nrm = torch.sqrt(torch.clamp((layer.V1 ** 2).sum(0), min=1e-12))
        # the conditional c1 is a 3x3 convolution (cond:427): V1 [9*C, F], rows ordered (kh, kw, cin)
        W1 = (layer.V1 * (layer.g1 / nrm)).view(3, 3, C, F).permute(3, 2, 0, 1)
        pre1 = Fn.conv2d((x * b).permute(0, 3, 1, 2), W1, padding=1).permute(0, 2, 3, 1).reshape(-1, F) + layer.b1
        tmp = alloc_slabs(1, N, H, W, x.device)[0]
        pres = []
        for l in range(2 * R):
            _lib.call('bruno_conv3x3', 4, slabs[l], wfwd[l], layer.br[l].contiguous(), 0, None, None, tmp, N, H, W)
            pres.append(slab_to_nhwc(tmp, N, H, W))
        for i in range(nl):
            _assign(layer.Lg[i * F:(i + 1) * F], layer.Lb[i * F:(i + 1) * F], h_pre[:, i * F:(i + 1) * F], (0,))
        _assign(layer.g1, layer.b1, pre1, (0,))
        for l in range(2 * R):
            _assign(layer.gr[l], layer.br[l], pres[l], (0, 1, 2))
        return y, ld_out


def cond_dense_coupling_init(layer, x, ld, y_label):
    from .nn_extra_nvp_conditional import _sgemm
    import torch.nn.functional as Fn
    with torch.no_grad():
        xs = x.shape
        N = xs[0]
        D = x.numel() // N
        out, ld_out = layer._apply(x, ld, y_label, False)        # layer output with the old g, b (builds the variables)
        P = layer.P
        idx = torch.arange(D, device=x.device) % 2
        b = (idx if layer.mask_type == 'even' else 1 - idx).to(torch.float32)
        y = x.reshape(N, D) * b
        todo = []
        for nm in ('d1', 'd2'):
            V, g, bb = P[nm + '/label_h/V'], P[nm + '/label_h/g'], P[nm + '/label_h/b']
            hp = _sgemm(y_label, V, V.shape[1]) * (g / torch.sqrt((V * V).sum(0))) + bb
            todo.append((g, bb, hp))
            o1 = torch.cat([Fn.leaky_relu(hp, 0.2), y], 1).contiguous()
            V, g, bb = P['%s/%s/V' % (nm, nm)], P['%s/%s/g' % (nm, nm)], P['%s/%s/b' % (nm, nm)]
            pre = _sgemm(o1, V, V.shape[1]) * (g / torch.sqrt((V * V).sum(0))) + bb
            todo.append((g, bb, pre))
            y = Fn.elu(pre)
        for g, bb, pre in todo:
            _assign(g, bb, pre, (0,))
        return out, ld_out
