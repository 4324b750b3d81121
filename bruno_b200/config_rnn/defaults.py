"""Test-time overrides of the config_rnn modules.

The reference keeps a handful of module-level knobs (sequence length, few-shot batch size, correlation mask) that its
test drivers overwrite from the command line BEFORE importing a config module (config_rnn/defaults.py, used at
test_nll.py:19,27).  Same names and semantics here; the assignment logic is table driven.
"""

# name -> value; every entry is also published as a module attribute (defaults.seq_len, ...)
_VALUES = {
    'seq_len': 20,
    'seq_len_few_shot': 2,
    'batch_size_few_shot': 20,
    'eval_only_last': False,
    'eps_corr': None,
    'mask_dims': False,
}

# command-line flag -> [(attribute, conversion)].  batch_size really is cast to bool in the reference (defaults.py:21);
# the few-shot configs only test its truthiness.
_FLAGS = {
    'seq_len': [('seq_len', None), ('seq_len_few_shot', None)],
    'batch_size': [('batch_size_few_shot', bool)],
    'eval_only_last': [('eval_only_last', bool)],
    'eps_corr': [('eps_corr', None)],
    'mask_dims': [('mask_dims', bool)],
}


def _publish():
    globals().update(_VALUES)


def set_parameters(args):
    """Copy the recognised fields of an argparse namespace into the module attributes."""
    given = vars(args)
    for flag, targets in _FLAGS.items():
        if flag not in given:
            continue
        for attr, conv in targets:
            _VALUES[attr] = conv(given[flag]) if conv is not None else given[flag]
    _publish()


_publish()
