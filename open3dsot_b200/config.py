"""YAML -> attribute dict, with the reference's override rule (main.py:23-49): every argparse value,
defaults included, overwrites the YAML key of the same name."""
import yaml

from .compat import EasyDict


def load_yaml(file_name):
    with open(file_name, "r") as f:
        return yaml.safe_load(f)


def load_config(cfg_path, overrides=None):
    cfg = load_yaml(cfg_path)
    if overrides:
        cfg.update(overrides)
    return EasyDict(cfg)
