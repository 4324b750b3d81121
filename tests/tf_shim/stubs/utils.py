"""Stub of the reference's `utils` for config import (`find_model_metadata` is read by the *_ft_* configs)."""


def find_model_metadata(metadata_dir, config_name):
    return '%s/%s-stub' % (metadata_dir.rstrip('/'), config_name)
