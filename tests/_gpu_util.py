"""Helpers shared by the GPU parity tests (not a test module)."""
import contextlib
import importlib

import torch

from oracle import bruno_oracle as O


def fresh_config(name):
    return importlib.reload(importlib.import_module('bruno_b200.config_rnn.' + name))


def load_model(name, P, x_example):
    """Fresh config module with its variables created (shaped by the first input) and loaded from the oracle dict P."""
    cfg = fresh_config(name)
    with torch.no_grad():
        cfg.build_model(x_example[:1].float().cuda())
    cfg.model.load_tf_variables(P)
    return cfg


def rel_err(a, b):
    """max |a - b| / max(|b|, 1): the per-example log-likelihood metric of the north star."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).abs() / b.abs().clamp_min(1.0)).max().item()


def grads_by_tf_name(cfg, loss):
    """{TF variable name: gradient (float64, CPU) with the TF shape} of `loss` through the hand-written backward."""
    params = cfg.model.parameters()
    grads = torch.autograd.grad(loss, params)
    by_id = {id(p): g for p, g in zip(params, grads)}
    out = {}
    for tfname, view in cfg.model.named_tf_variables():
        base = view._base if view._base is not None else view
        g = by_id[id(base)] if id(base) in by_id else by_id[id(view)]
        out[tfname] = g.as_strided(view.shape, view.stride(), view.storage_offset() - base.storage_offset()).double().cpu()
    return out


@contextlib.contextmanager
def oracle_device(device):
    """Runs the oracle's torch code on `device`: its restatement is device-agnostic except for the factory calls
    (masks, zeros), which follow torch's default device.  On the GPU box the float64 oracle at the BENCHMARK shape
    (640 images, 2.8 TFLOP forward) takes seconds instead of minutes; it remains the checker, never the product."""
    with torch.device(device):
        yield


def params_to(P, device, dtype=torch.float64):
    return type(P)((k, v.to(device=device, dtype=dtype)) for k, v in P.items())


def oracle_forward(spec, P, x, device='cuda'):
    """(log_probs, latent, prior, z_vec) of the float64 oracle, evaluated on `device`, returned on the CPU."""
    with torch.no_grad(), oracle_device(device):
        out = O.build_model(spec, params_to(P, device), x.to(device=device, dtype=torch.float64))
    return [t.cpu() for t in out]


def sample_index(n, nsample=64):
    """The strided sample of a large gradient stored in tests/golden/ref_flow_*.npz (tools/make_ref_flow_golden.py)."""
    return torch.linspace(0, n - 1, nsample).round().long()
