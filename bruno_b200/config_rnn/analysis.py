"""The reference's analysis entry points that consume outputs 1-3 of build_model (latent log-probs, prior log-probs,
latent codes) and the process parameters -- config_rnn/test_anomaly.py, test_corr.py, test_latent_dist.py,
test_probs_evolution.py.  Same flags, same printed quantities; the reference renders matplotlib figures (absent here,
plotting is out of scope), these write the plotted arrays as .npy / .json next to the checkpoint instead.
All model evaluations go through config.build_model on the B200 path."""
import importlib
import json
import os

import numpy as np
import torch

from .. import utils
from . import defaults


def _load(config_name, metadata_dir):
    device = torch.device('cuda', 0)
    config = importlib.import_module('bruno_b200.config_rnn.%s' % config_name)
    save_dir = utils.find_model_metadata(metadata_dir + '/', config_name)
    print('exp_id', os.path.dirname(save_dir))
    utils.load_checkpoint_into(config, save_dir, device)
    return config, save_dir, device


def process_parameters(layer):
    """(mu, var, corr, nu) flattened; nu = 0 for the Gaussian process (test_corr.py:49-56, test_latent_dist.py:76-84)."""
    var = layer.var.detach().cpu().numpy().flatten()
    corr = layer.corr.detach().cpu().numpy().flatten()
    mu = layer.mu.detach().cpu().numpy().flatten()
    nu = layer.nu.detach().cpu().numpy().flatten() if getattr(layer, 'nu', None) is not None else np.zeros_like(var)
    return mu, var, corr, nu


def anomaly(config_name, n_sequences, metadata_dir='metadata', n_ignore=2, mask_dims=0):
    """test_anomaly.py:25-62: score_i = latent_log_prob_i - latent_log_prob_under_prior_i along ONE sequence; anomalies =
    elements (after the first n_ignore) whose score is below Q1 - 1.5 IQR."""
    config, save_dir, device = _load(config_name, metadata_dir)
    target = os.path.join(save_dir, 'anomaly')
    utils.autodir(target)
    data_iter = config.test_data_iter
    data_iter.batch_size = 1
    out = []
    for iteration, (x_batch, y_batch) in zip(range(n_sequences), data_iter.generate_anomaly()):
        assert x_batch.shape[0] == 1
        with torch.no_grad():
            _, lp, lpp, _ = config.build_model(torch.from_numpy(x_batch).to(device))
        lp, lpp = lp.cpu().numpy().squeeze(), lpp.cpu().numpy().squeeze()
        scores = lp - lpp
        for i in range(y_batch.shape[-1]):
            print(i, lp[i], lpp[i], scores[i], y_batch[0, i])
        print('-------------')
        q1, q3 = np.percentile(scores[n_ignore:], [25, 75])
        lower_bound = q1 - (q3 - q1) * 1.5
        idxs = np.where(scores < lower_bound)[0]
        idxs = idxs[idxs >= n_ignore]                      # don't count the first images as outliers
        print(idxs)
        np.save(os.path.join(target, 'anomaly_%s_%s.npy' % (iteration, mask_dims)), scores)
        out.append((scores, idxs, y_batch[0]))
    return out


def corr_report(config_name, metadata_dir='metadata'):
    """test_corr.py:49-81: the learned nu / var / corr and the number of latent dimensions with corr > eps."""
    config, save_dir, device = _load(config_name, metadata_dir)
    target = os.path.join(save_dir, 'misc_plots')
    utils.autodir(target)
    mu, var, corr, nu = process_parameters(config.student_layer)
    print('nu\n', nu)
    print('--------------------')
    print('var\n', var)
    print('--------------------')
    print('corr\n', corr)
    print('******* corr - nu ********')
    for i in np.where(corr > 0.05)[0]:
        print(i, corr[i], nu[i], var[i])
    print('--------------------------')
    print('VAR', np.min(var), np.max(var))
    eps_range, n_remained = [], []
    for t in np.arange(0, 0.9, 0.001):
        nr = int(np.sum(corr > t))
        if not n_remained or nr != n_remained[-1]:
            n_remained.append(nr)
            eps_range.append(float(t))
        if t < 0.1:
            print(t, n_remained[-1])
    with open(os.path.join(target, 'eps_plot.json'), 'w') as f:
        json.dump({'eps': eps_range, 'n_dims': n_remained}, f)
    np.save(os.path.join(target, 'nu_corr.npy'), np.stack([nu, corr]))
    return eps_range, n_remained


def student_pdf_1d(X, mu, var, nu):                       # test_latent_dist.py:28-35
    import math
    if nu > 50:
        return gauss_pdf_1d(X, mu, var)
    num = math.gamma((1. + nu) / 2.) * np.power(1. + (1. / (nu - 2)) * (1. / var * (X - mu) ** 2), -(1. + nu) / 2.)
    return num / (math.gamma(nu / 2.) * np.power((nu - 2) * math.pi * var, 0.5))


def gauss_pdf_1d(X, mu, var):                             # test_latent_dist.py:38-39
    return 1. / np.sqrt(2. * np.pi * var) * np.exp(- (X - mu) ** 2 / (2. * var))


def latent_dist(config_name, set_name='train', metadata_dir='metadata', batch_size=1000, n_bins=100):
    """test_latent_dist.py:60-120: per latent dimension, the histogram of the codes of the FIRST element of `batch_size`
    sequences next to the prior marginal (Student-t, or Gaussian for *gp* configs / nu > 50)."""
    config, save_dir, device = _load(config_name, metadata_dir)
    target = os.path.join(save_dir, 'hists_%s' % set_name)
    utils.autodir(target)
    gp_model = 'gp' in config_name
    data_iter = config.test_data_iter if set_name == 'test' else config.train_data_iter
    data_iter.batch_size = batch_size
    mu, var, corr, nu = process_parameters(config.student_layer)
    x_batch = next(data_iter.generate())
    with torch.no_grad():
        codes = config.build_model(torch.from_numpy(x_batch).to(device), want_prior=False)[3]
    all_codes = codes[:, 0, :].cpu().numpy()               # only the first element of each sequence
    print(all_codes.shape)
    hists = np.zeros((all_codes.shape[-1], n_bins), dtype=np.float32)
    pdfs = np.zeros((all_codes.shape[-1], n_bins), dtype=np.float32)
    lims = np.zeros((all_codes.shape[-1], 2), dtype=np.float32)
    for i in range(all_codes.shape[-1]):
        lo = min(mu[i] - 5 * np.sqrt(var[i]), np.min(all_codes[:, i]))
        hi = max(mu[i] + 5 * np.sqrt(var[i]), np.max(all_codes[:, i]))
        centres = np.linspace(lo, hi, n_bins)
        pdfs[i] = gauss_pdf_1d(centres, mu[i], var[i]) if gp_model else student_pdf_1d(centres, mu[i], var[i], nu[i])
        hists[i] = np.histogram(all_codes[:, i], bins=n_bins, range=(lo, hi), density=True)[0]
        lims[i] = (lo, hi)
        if i < 10:
            print(i, np.min(all_codes[:, i]), np.max(all_codes[:, i]), np.argmin(all_codes[:, i]), np.argmax(all_codes[:, i]))
    np.savez(os.path.join(target, 'hists.npz'), hists=hists, pdfs=pdfs, lims=lims, corr=corr, nu=nu, var=var)
    return hists, pdfs


def probs_evolution(config_name, set_name='test', n_batches=1000, same_class=1, same_image=0, black_image=0,
                    metadata_dir='metadata'):
    """test_probs_evolution.py:60-105: feed seq_len rolled copies of one sequence; the diagonal of the [seq_len, seq_len]
    outputs follows frame 0 through every position: score_n = latent log-prob - prior log-prob after n observations."""
    config, save_dir, device = _load(config_name, metadata_dir)
    target = os.path.join(save_dir, 'misc_plots')
    utils.autodir(target)
    data_iter = config.test_data_iter if set_name == 'test' else config.train_data_iter
    rng = np.random.RandomState(42)
    scores, prior_ll, probs_x = [], [], []
    gen = data_iter.generate_diagonal_roll(same_class=same_class, noise_rng=rng, same_image=same_image, black_image=black_image)
    for idx, x_batch in zip(range(n_batches), gen):
        with torch.no_grad():
            lp_x, lp_z, lp_prior_z, _ = config.build_model(torch.from_numpy(x_batch).to(device))
        lp_x, lp_z, lp_prior_z = lp_x.cpu().numpy(), lp_z.cpu().numpy(), lp_prior_z.cpu().numpy()
        scores.append(np.diag(lp_z - lp_prior_z))
        prior_ll.append(lp_x[0, 0])
        probs_x.append(np.diag(lp_x))
    scores_mean = np.mean(np.asarray(scores), axis=0)
    print('avg LL under prior :', np.mean(prior_ll))
    ll = np.mean(np.asarray(probs_x), axis=0)
    print('avg LL:', ll)
    tag = 'len%s_%s_class%s_img%s_black%s' % (config.seq_len, set_name, same_class, same_image, black_image)
    np.save(os.path.join(target, 'scores_plot_%s.npy' % tag), scores_mean)
    np.save(os.path.join(target, 'll_plot_%s.npy' % tag), ll)
    return scores_mean, ll
