"""NLL / bits-per-dim of Conditional BRUNO -- counterpart of the reference's config_conditional/test_nll.py (same flags
--config_name --seq_len --n_context --n_batches and prints); restores the torch checkpoint written by train.py."""
import argparse
import importlib
import json
import os

import numpy as np
import torch

from .. import utils
from . import defaults


def restore(config, save_dir, device):
    ck = torch.load(os.path.join(save_dir, 'params.pt'), map_location='cpu', weights_only=False)
    it = config.test_data_iter
    x, y = next(it.generate())
    config.build_model(torch.from_numpy(x[:1]).to(device), torch.from_numpy(y[:1]).to(device))    # creates the variables
    config.load_tf_variables(ck['variables'])
    return ck


def evaluate(config, data_iter, n_batches, device):
    data_iter.batch_size = 1                                               # test_nll.py:57
    probs = []
    for _, (x_batch, y_batch) in zip(range(0, n_batches), data_iter.generate()):
        with torch.no_grad():
            l = config.build_model(torch.from_numpy(x_batch).to(device), torch.from_numpy(y_batch).to(device))[0]
        probs.append(l.cpu().numpy())
    avg_loss = -1. * np.mean(probs)
    return avg_loss, avg_loss / np.log(2.) / config.ndim


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--seq_len', type=int, default=16, help='Sequence length')
    parser.add_argument('--n_context', type=int, default=1, help='Context length')
    parser.add_argument('--n_batches', type=int, default=1000, help='Number of batches')
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args, _ = parser.parse_known_args(argv)
    defaults.set_parameters(args)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    device = torch.device('cuda', 0)
    config = importlib.reload(importlib.import_module('bruno_b200.config_conditional.%s' % args.config_name))
    save_dir = utils.find_model_metadata(args.metadata_dir + '/', args.config_name)
    print('exp_id', os.path.dirname(save_dir))
    print('restoring parameters from', save_dir + 'params.pt')
    restore(config, save_dir, device)
    print(args.config_name)
    print('Sequence length:', config.seq_len)
    print('Context length:', config.n_context)
    test = evaluate(config, config.test_data_iter, args.n_batches, device)
    print('Test Loss %.3f' % test[0])
    print('Bits per dim %.3f' % test[1])
    train = evaluate(config, config.train_data_iter, args.n_batches, device)
    print('Train Loss %.3f' % train[0])
    print('Bits per dim %.3f' % train[1])
    return test, train


if __name__ == '__main__':
    main()
