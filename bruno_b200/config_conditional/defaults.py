"""Test-time overrides of the config_conditional modules (reference: config_conditional/defaults.py): the test drivers
set the sequence and context lengths from the command line before the config module is imported."""

_OVERRIDABLE = ('seq_len', 'n_context')

seq_len = 16       # frames per sequence
n_context = 1      # frames the Gaussian process is conditioned on before scoring / sampling


def set_parameters(args):
    """Take `seq_len` / `n_context` from an argparse namespace when present."""
    given = vars(args)
    module = globals()
    for key in _OVERRIDABLE:
        if key in given:
            module[key] = given[key]
