"""Model registry.  Mirror of models/__init__.py:19-21: yaml `net_model: BAT | P2B | m2track` -> class."""
from . import bat, p2b  # noqa: F401

try:  # M2-Track is a dense-MLP model (no pointnet2 ops); optional while it is being widened
    from . import m2track  # noqa: F401
except ImportError:  # pragma: no cover
    m2track = None


def get_model(name):
    return globals()[name.lower()].__getattribute__(name.upper())
