"""Shared implementation of the config builders' `build_model` (config_rnn/bn2_omniglot_tp.py:59-168,
en2_noconv_mnist_tp.py:57-160, an2_cifar_tp.py:44-153) on the B200 layer mirrors, plus parameter bookkeeping:
one flat fp32 parameter buffer + one flat gradient buffer (what the optimiser kernel and the NCCL allreduce see),
and import/export under the reference's TF variable names."""
from collections import OrderedDict

import numpy as np
import torch

from . import _lib, nn_extra_nvp


class BrunoModel(object):
    """Holds the layer lists a reference config module keeps as globals (nvp_layers, nvp_dense_layers,
    student_layer) and runs the forward / sampling graph of build_model."""

    def __init__(self, nvp_layers, nvp_dense_layers, student_layer, n_samples=4):
        self.nvp_layers = nvp_layers
        self.nvp_dense_layers = nvp_dense_layers
        self.student_layer = student_layer
        self.n_samples = n_samples
        self.flat = None
        self.flat_grad = None
        self._index = None

    # ---------------------------------------------------------------------------------------------------
    # forward graph
    # ---------------------------------------------------------------------------------------------------
    def flow_forward(self, x_bs):
        """x_bs [N,H,W,C] pixels in [0,256) -> (z [N,h,w,c], log_det_jac [N]);  bn2:74-96."""
        y, log_det_jac = nn_extra_nvp.preprocess_forward(x_bs)              # ScaleLayer + LogitLayer, bn2:82-83
        z = None
        for layer in self.nvp_layers:                                        # bn2:87-88
            y, z, log_det_jac = layer.forward_and_jacobian(y, z, log_det_jac)
        if self.nvp_layers:
            z = nn_extra_nvp.channel_concat(z, y)                            # bn2:90
        else:
            z = y
        for layer in self.nvp_dense_layers:                                  # bn2:91-92
            z, _, log_det_jac = layer.forward_and_jacobian(z, None, log_det_jac)
        return z, log_det_jac

    def flow_inverse(self, z_img):
        """z_img [N,h,w,c] -> (x [N,H,W,C], log_det_jac [N]) with log_det starting at 0;  bn2:139-149."""
        log_det_jac = torch.zeros(z_img.shape[0], device=z_img.device)
        z = z_img
        for layer in reversed(self.nvp_dense_layers):
            z, _, log_det_jac = layer.backward(z, None, log_det_jac)
        if self.nvp_layers:
            x = None
            for layer in reversed(self.nvp_layers):
                x, z, log_det_jac = layer.backward(x, z, log_det_jac)
        else:
            x = z
        return nn_extra_nvp.preprocess_backward(x, log_det_jac)

    def build_model(self, x, init=False, sampling_mode=False, mask_dim=None, noise=None, want_prior=True):
        with nn_extra_nvp.init_scope(init):
            x_shape = nn_extra_nvp.int_shape(x)
            B, T = x_shape[0], x_shape[1]
            x_bs = x.reshape(B * T, x_shape[2], x_shape[3], x_shape[4])
            z, log_det_jac = self.flow_forward(x_bs)
            z_shape = nn_extra_nvp.int_shape(z)
            z_vec = z.reshape(B, T, -1)
            log_det_jac = log_det_jac.reshape(B, T)
            if sampling_mode:
                return self._sample(z_vec, z_shape, x_shape, noise)
            latent_log_probs, latent_log_probs_prior = self.student_layer.sequence_log_likelihoods(
                z_vec, mask_dim=mask_dim, want_prior=want_prior)             # the unrolled loop of bn2:103-124
            log_probs = latent_log_probs + log_det_jac                       # bn2:116
            return log_probs, latent_log_probs, latent_log_probs_prior, z_vec

    def _sample(self, z_vec, z_shape, x_shape, noise):
        """sampling_mode=True branch, bn2:106-158: a sample from the posterior predictive BEFORE each observation and
        one more after the last, pushed through the inverse flow.  noise[i] (optional) = the random draws of step i."""
        sl = self.student_layer
        with torch.no_grad():
            B, T, D = z_vec.shape
            n = self.n_samples
            mu, var, nu = sl.posterior_sequence(z_vec)                       # after 0..T observations
            xz = z_vec - sl.mu.view(1, 1, D)
            s1 = torch.cat([torch.zeros(B, 1, D, device=z_vec.device), torch.cumsum(xz, 1)], 1)
            s2 = torch.cat([torch.zeros(B, 1, D, device=z_vec.device), torch.cumsum(xz * xz, 1)], 1)
            z_samples, latent = [], []
            for i in range(T + 1):
                Bi = 1 if i == 0 else B                                      # the prior has batch dim 1 (student:30)
                nz = noise[i] if noise is not None else sl.draw_noise(n, Bi, z_vec.device)
                zs = sl.sample_from(mu[:Bi, i], var[:Bi, i], None if nu is None else nu[i], nz)   # [n,Bi,D]
                z_samples.append(zs)
                # latent_log_prob = get_log_likelihood(z_sample[:, 0, :]) under the current distribution (bn2:110)
                obs = zs[:, 0, :].contiguous()
                nb = obs.shape[0]
                rows = (torch.arange(nb, device=obs.device) % Bi) if Bi > 1 else torch.zeros(nb, dtype=torch.long, device=obs.device)
                latent.append(sl.log_likelihood_with_state(obs, s1[rows, i].contiguous(), s2[rows, i].contiguous(), i))
            z_samples = torch.cat(z_samples, 1)                              # [n, sum(Bi), D]
            zs_shape = z_samples.shape
            z_img = z_samples.reshape(zs_shape[0] * zs_shape[1], z_shape[1], z_shape[2], z_shape[3]).contiguous()
            x_samples, ld = self.flow_inverse(z_img)
            x_samples = x_samples.reshape(zs_shape[0], zs_shape[1], x_shape[2], x_shape[3], x_shape[4])
            ld = ld.reshape(zs_shape[0], zs_shape[1])
            latent = torch.stack(latent, 1)                                  # [n, T+1]
            log_probs = latent - ld[:, :T + 1]                               # bn2:153-156
            return x_samples, log_probs

    # ---------------------------------------------------------------------------------------------------
    # CUDA-graphed inference (latency-bound cases: few-shot episodes of 40 images, test_nll's batch of 1)
    # ---------------------------------------------------------------------------------------------------
    def graphed_forward(self, x_example, want_prior=False):
        """Captures build_model(x) (forward, no grad) for the shape of `x_example` into ONE CUDA graph and returns
        f(x) -> (log_probs, latent_log_probs, latent_log_probs_prior or None, z_vec).  ~450 kernel launches of a
        20-way 1-shot episode become one graph launch; outputs are views of static buffers (overwritten by the next call)."""
        static_x = x_example.clone()
        with torch.no_grad():
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):                       # warm-up: lazy variable creation, cudaFuncSetAttribute, slab pools
                    self.build_model(static_x, want_prior=want_prior)
            torch.cuda.current_stream().wait_stream(s)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):             # the warmed-up stream: its GEMM workspace is bound
                static_out = self.build_model(static_x, want_prior=want_prior)

        def run(x):
            static_x.copy_(x, non_blocking=True)
            graph.replay()
            return static_out
        run.graph = graph
        return run

    # ---------------------------------------------------------------------------------------------------
    # parameters
    # ---------------------------------------------------------------------------------------------------
    def layers(self):
        return list(self.nvp_layers) + list(self.nvp_dense_layers)

    def named_parameters(self):
        out = []
        for layer in self.layers():
            if hasattr(layer, 'named_parameters'):
                out += layer.named_parameters()
        out += self.student_layer.named_parameters()
        return out

    def parameters(self):
        return [p for _, p in self.named_parameters()]

    def materialize(self, x_example):
        """Creates the variables (they are shaped by the first input, like tf.get_variable under make_template)."""
        with torch.no_grad():
            self.build_model(x_example, want_prior=False)
        return self

    def flatten_parameters(self):
        """Re-homes every parameter into ONE flat fp32 buffer (and gives each a .grad view into ONE flat gradient
        buffer): a single optimiser launch and a single NCCL allreduce per step."""
        named = self.named_parameters()
        sizes = [(p.numel() + 3) // 4 * 4 for _, p in named]                 # keep every tensor 16-byte aligned
        total = int(sum(sizes))
        dev = named[0][1].device
        flat = torch.zeros(total, device=dev)
        grad = torch.zeros(total, device=dev)
        index = OrderedDict()
        off = 0
        new = {}
        for (name, p), sz in zip(named, sizes):
            view = flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.detach())
            view.requires_grad_(True)
            view.grad = grad[off:off + p.numel()].view(p.shape)
            index[name] = (off, p.numel(), tuple(p.shape))
            new[name] = view
            off += sz
        for layer in self.layers():
            if hasattr(layer, '_set_parameters'):
                layer._set_parameters([new[n] for n, _ in layer.named_parameters()])
        sl = self.student_layer
        for n, _ in sl.named_parameters():
            attr = {'prior_var': 'var_vbl', 'prior_nu': 'nu_vbl', 'prior_corr': 'corr_vbl'}[n.split('/')[-1]]
            setattr(sl, attr, new[n])
        self.flat, self.flat_grad, self._index = flat, grad, index
        return flat, grad

    def student_param_ranges(self):
        """[(offset, numel)] of the variables whose gradient train.py:108-111 rescales."""
        return [(o, n) for k, (o, n, _) in self._index.items() if 'prior_' in k]

    def named_tf_variables(self, prefix='model/'):
        out = []
        for layer in self.layers():
            out += layer.named_tf_variables(prefix)
        sl = self.student_layer
        out += [(prefix + n, p) for n, p in sl.named_parameters()]
        return out

    def load_tf_variables(self, params):
        """params: {TF variable name: array-like} (e.g. the oracle's parameter dict or a converted checkpoint)."""
        with torch.no_grad():
            for name, view in self.named_tf_variables():
                src = torch.as_tensor(np.asarray(params[name]) if not torch.is_tensor(params[name]) else params[name])
                view.copy_(src.to(torch.float32).reshape(view.shape).to(view.device))

    def export_tf_variables(self):
        return OrderedDict((n, v.detach().cpu().numpy().copy()) for n, v in self.named_tf_variables())
