"""Stand-ins for the reference's optional host-side dependencies that are absent from this image
(easydict, pytorch_lightning): used only when the real package cannot be imported."""
from .easydict import EasyDict  # noqa: F401
from .lightning import LightningModule  # noqa: F401
