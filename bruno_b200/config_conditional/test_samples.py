"""Conditional sampling -- counterpart of the reference's config_conditional/test_samples.py (flags --config_name --set
--n_samples --seq_len --n_context): for each test sequence, condition the GP on the first n_context frames and sample
every frame given its view label; the images are saved as .npy (the reference plots them with matplotlib, which is
out of scope) next to the checkpoint."""
import argparse
import importlib
import json
import os

import numpy as np
import torch

from .. import utils
from . import defaults
from .test_nll import restore


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--set', type=str, default='test', help='Test or train part')
    parser.add_argument('--n_samples', type=int, default=3, help='Number of sequences')
    parser.add_argument('--seq_len', type=int, default=20, help='Sequence length')
    parser.add_argument('--n_context', type=int, default=1, help='Context length')
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args, _ = parser.parse_known_args(argv)
    defaults.set_parameters(args)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    device = torch.device('cuda', 0)
    config = importlib.reload(importlib.import_module('bruno_b200.config_conditional.%s' % args.config_name))
    save_dir = utils.find_model_metadata(args.metadata_dir + '/', args.config_name)
    print('exp_id', os.path.dirname(save_dir))
    print('n_context', config.n_context)
    print('seq_len', config.seq_len)
    print('restoring parameters from', save_dir + 'params.pt')
    restore(config, save_dir, device)
    data_iter = config.test_data_iter if args.set == 'test' else config.train_data_iter
    data_iter.batch_size = 1
    out = []
    for i, (x_batch, y_batch) in zip(range(args.n_samples), data_iter.generate()):
        print(i, "Generating samples...")
        with torch.no_grad():
            xs = config.build_model(torch.from_numpy(x_batch).to(device), torch.from_numpy(y_batch).to(device), sampling_mode=True)
        out.append({'x': x_batch, 'y_label': y_batch, 'samples': xs.cpu().numpy()})
        assert np.isfinite(out[-1]['samples']).all()
    target = os.path.join(save_dir, 'samples_%s.npy' % args.set)
    np.save(target, np.array(out, dtype=object), allow_pickle=True)
    print('Saved to', target)
    return out


if __name__ == '__main__':
    main()
