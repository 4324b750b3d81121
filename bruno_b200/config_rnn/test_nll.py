"""NLL / bits-per-dim evaluation -- counterpart of the reference's config_rnn/test_nll.py (same flags and prints)."""
import argparse
import importlib
import json
import os

import numpy as np
import torch

from .. import utils
from . import defaults


def evaluate(config, data_iter, n_batches, device):
    data_iter.batch_size = 1                                               # test_nll.py:53
    probs, probs_prior = [], []
    for _, x_batch in zip(range(0, n_batches), data_iter.generate()):
        with torch.no_grad():
            l = config.build_model(torch.from_numpy(x_batch).to(device), want_prior=False)[0].cpu().numpy()
        probs.append(l)
        probs_prior.append(l[:, 0])
    avg_loss = -1. * np.mean(probs)
    bits_per_dim = avg_loss / np.log(2.) / config.ndim
    avg_loss_prior = -np.mean(probs_prior)
    bits_per_dim_prior = avg_loss_prior / np.log(2.) / config.ndim
    return avg_loss, bits_per_dim, avg_loss_prior, bits_per_dim_prior


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--seq_len', type=int, default=20, help='sequence length')
    parser.add_argument('--n_batches', type=int, default=10000, help='number of batches')
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args, _ = parser.parse_known_args(argv)
    defaults.set_parameters(args)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    device = torch.device('cuda', 0)
    config = importlib.import_module('bruno_b200.config_rnn.%s' % args.config_name)
    save_dir = utils.find_model_metadata(args.metadata_dir + '/', args.config_name)
    print('exp_id', os.path.dirname(save_dir))
    print('restoring parameters from', save_dir + 'params.pt')
    utils.load_checkpoint_into(config, save_dir, device)
    print('Sequence length:', args.seq_len)
    test_losses = evaluate(config, config.test_data_iter, args.n_batches, device)
    print('Test Loss %.3f' % test_losses[0])
    print('Bits per dim %.3f' % test_losses[1])
    print('Test loss under prior %.3f' % test_losses[2])
    train_losses = evaluate(config, config.train_data_iter, args.n_batches, device)
    print('Train Loss %.3f' % train_losses[0])
    print('Bits per dim %.3f' % train_losses[1])
    print('Train loss under prior %.3f' % train_losses[2])
    print('%.3f(%.3f) | %.3f(%.3f) | %.3f(%.3f) | %.3f(%.3f)' % (test_losses + train_losses))
    return test_losses, train_losses


if __name__ == '__main__':
    main()
