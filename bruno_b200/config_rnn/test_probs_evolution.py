"""How the predictive score of a frame evolves with the number of observations -- counterpart of the reference's
config_rnn/test_probs_evolution.py (same flags)."""
import argparse
import json

from . import defaults


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--set', type=str, default='test', help='train or test set?')
    parser.add_argument('--seq_len', type=int, default=100, help='sequence length')
    parser.add_argument('--mask_dims', type=int, default=0, help='keep the dimensions with correlation > eps_corr?')
    parser.add_argument('--eps_corr', type=float, default=0., help='minimum correlation')
    parser.add_argument('--same_class', type=int, default=1, help='sequences from the same class')
    parser.add_argument('--same_image', type=int, default=0, help='sequences from the same image')
    parser.add_argument('--black_image', type=int, default=0, help='black image on diagonal')
    parser.add_argument('--n_batches', type=int, default=1000, help='how many batches to average over')
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args, _ = parser.parse_known_args(argv)
    defaults.set_parameters(args)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    if args.mask_dims == 0:
        assert args.eps_corr == 0.
    from . import analysis
    return analysis.probs_evolution(args.config_name, args.set, args.n_batches, args.same_class, args.same_image,
                                    args.black_image, args.metadata_dir)


if __name__ == '__main__':
    main()
