"""Histograms of the latent codes against the prior marginals -- counterpart of the reference's
config_rnn/test_latent_dist.py (same flags)."""
import argparse
import json

from . import defaults


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='name of the configuration')
    parser.add_argument('--seq_len', type=int, default=2, help='sequence length')
    parser.add_argument('--set', type=str, default='train', help='train or test set?')
    parser.add_argument('--batch_size', type=int, default=1000)
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args, _ = parser.parse_known_args(argv)
    bs = args.batch_size
    del args.batch_size                                    # not the few-shot batch size of defaults.set_parameters
    defaults.set_parameters(args)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    from . import analysis
    return analysis.latent_dist(args.config_name, args.set, args.metadata_dir, batch_size=bs)


if __name__ == '__main__':
    main()
