"""autodir / find_model_metadata with the reference's conventions (utils.py:12-24) + the stdout tee (logger.py:4-14)."""
import glob
import os
import sys


def autodir(path):
    if not os.path.isdir(path):
        os.makedirs(path)


def find_model_metadata(metadata_dir, config_name):
    metadata_paths = glob.glob(metadata_dir + '/%s-*/' % config_name)
    print(metadata_dir, config_name, metadata_paths)
    if not metadata_paths:
        raise ValueError('No metadata files for config %s' % config_name)
    elif len(metadata_paths) > 1:
        raise ValueError('Multiple metadata files for config %s' % config_name)
    return metadata_paths[0]


class Logger(object):
    def __init__(self, path):
        self.terminal = sys.stdout
        self.log = open(path, 'a')

    def write(self, message):
        self.terminal.write(message)
        self.log.write(message)

    def flush(self):
        self.terminal.flush()
        self.log.flush()


def load_checkpoint_into(config, save_dir, device):
    """Restores params.pt (written by train.py) into config.model; the variables are created first from a dummy batch."""
    import torch
    ck = torch.load(os.path.join(save_dir, 'params.pt'), map_location='cpu', weights_only=False)
    x0 = torch.full((1,) + tuple(config.obs_shape), 128.0, device=device)
    with torch.no_grad():
        config.build_model(x0, want_prior=False)
    config.model.load_tf_variables(ck['variables'])
    return ck
