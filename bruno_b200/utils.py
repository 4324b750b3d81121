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
    """Restores a checkpoint into config.model; the variables are created first from a dummy batch.  Two formats:
    params.pt (written by this package's train.py: variables under the reference's TF names + optimiser slots) and a
    checkpoint written by the TF-1.7 REFERENCE itself (`saver.save(sess, save_dir + '/params.ckpt')`, train.py:232 -- a V2
    tensor bundle, read by bruno_b200/tf_checkpoint.py without TensorFlow); params.pt wins when both exist."""
    import torch
    x0 = torch.full((1,) + tuple(config.obs_shape), 128.0, device=device)
    with torch.no_grad():
        config.build_model(x0, want_prior=False)
    pt = os.path.join(save_dir, 'params.pt')
    if os.path.exists(pt):
        ck = torch.load(pt, map_location='cpu', weights_only=False)
        config.model.load_tf_variables(ck['variables'])
        return ck
    from . import tf_checkpoint
    prefix = tf_checkpoint.latest_checkpoint(save_dir)
    if prefix is None:
        raise ValueError('no params.pt and no TensorFlow checkpoint in %s' % save_dir)
    print('restoring parameters from', prefix)
    wanted = set(n for n, _ in config.model.named_tf_variables())
    variables = tf_checkpoint.read_checkpoint(prefix, names=wanted)
    missing = sorted(wanted - set(variables))
    if missing:
        raise ValueError('%s lacks %d variables of this model, e.g. %s' % (prefix, len(missing), missing[:3]))
    config.model.load_tf_variables(variables)
    return {'variables': variables}


def export_tf_checkpoint(config, save_dir, name='params.ckpt'):
    """Writes config.model's variables as a TensorFlow V2 checkpoint the reference's `saver.restore` can read."""
    from . import tf_checkpoint
    autodir(save_dir)
    return tf_checkpoint.write_checkpoint(os.path.join(save_dir, name), config.model.export_tf_variables())
