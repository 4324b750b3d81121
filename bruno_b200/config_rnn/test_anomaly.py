"""Outlier detection along one sequence -- counterpart of the reference's config_rnn/test_anomaly.py (same flags)."""
import argparse
import json

from . import defaults


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='name of the configuration')
    parser.add_argument('--seq_len', type=int, default=100, help='sequence length')
    parser.add_argument('--n_sequences', type=int, default=1, help='number of sequences')
    parser.add_argument('--mask_dims', type=int, default=0, help='keep the dimensions with correlation > eps_corr')
    parser.add_argument('--eps_corr', type=float, default=0., help='minimum correlation')
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args, _ = parser.parse_known_args(argv)
    defaults.set_parameters(args)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    if args.mask_dims == 0:
        assert args.eps_corr == 0.
    from . import analysis
    return analysis.anomaly(args.config_name, args.n_sequences, args.metadata_dir, mask_dims=args.mask_dims)


if __name__ == '__main__':
    main()
