"""Model registry.  Mirror of models/__init__.py:19-21: yaml `net_model: BAT | P2B | m2track` -> class."""
from . import bat, m2track, p2b  # noqa: F401


def get_model(name):
    return globals()[name.lower()].__getattribute__(name.upper())
