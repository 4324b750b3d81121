"""Learned process parameters and the corr > eps dimension count -- counterpart of the reference's config_rnn/test_corr.py."""
import argparse
import json


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args, _ = parser.parse_known_args(argv)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    from . import analysis
    return analysis.corr_report(args.config_name, args.metadata_dir)


if __name__ == '__main__':
    main()
