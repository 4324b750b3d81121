"""Imports the UNMODIFIED reference modules from /root/reference over the eager `tensorflow` stand-in.

TEST INFRASTRUCTURE ONLY (see tensorflow/__init__.py).  /root/reference exists only in the build container, so
everything here is used by (a) tools/make_ref_flow_golden.py, which writes the committed fixtures, and (b) CPU tests
that skip when the reference tree is absent.
"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get('BRUNO_REFERENCE', '/root/reference')
_REF_MODULES = ('nn_extra_nvp', 'nn_extra_student', 'nn_extra_gauss', 'nn_extra_nvp_conditional', 'data_iter',
                'data_iter_conditional', 'utils', 'utils_conditional', 'config_rnn', 'config_conditional', 'tensorflow')


def available():
    return os.path.isfile(os.path.join(REFERENCE, 'nn_extra_nvp.py'))


class reference_modules(object):
    """Context: sys.path / sys.modules set up so that `import tensorflow`, `import data_iter`, `import nn_extra_nvp`,
    `from config_rnn import defaults` resolve to the shim, the stubs and the reference tree; restored on exit so the
    repo's own same-named packages (bruno_b200.config_rnn ...) and any real module are untouched."""

    def __enter__(self):
        assert available(), 'reference tree not found at %s' % REFERENCE
        self.saved_path = list(sys.path)
        self.saved_mods = {k: v for k, v in sys.modules.items() if k.split('.')[0] in _REF_MODULES}
        for k in self.saved_mods:
            del sys.modules[k]
        sys.path[:0] = [HERE, os.path.join(HERE, 'stubs'), REFERENCE]
        import tensorflow as tf
        assert tf.__file__.startswith(HERE)
        tf.reset_default_graph()
        return tf

    def __exit__(self, *a):
        for k in [k for k in sys.modules if k.split('.')[0] in _REF_MODULES]:
            del sys.modules[k]
        sys.modules.update(self.saved_mods)
        sys.path[:] = self.saved_path


def import_config(package, name, seq_len=None, **defaults_overrides):
    """Inside a `reference_modules()` context: import `config_rnn.<name>` / `config_conditional.<name>` fresh, with
    `defaults.seq_len` (and other defaults attributes) set first, the way the drivers' `set_parameters` does."""
    defaults = importlib.import_module(package + '.defaults')
    if seq_len is not None:
        defaults.seq_len = seq_len
        if hasattr(defaults, 'seq_len_few_shot'):
            defaults.seq_len_few_shot = seq_len
    for k, v in defaults_overrides.items():
        setattr(defaults, k, v)
    return importlib.import_module(package + '.' + name)
