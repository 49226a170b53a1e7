# same exports as /root/reference/Geom3D/models/__init__.py:1-2
from .painn import PaiNN
from .schnet import SchNet

__all__ = ["PaiNN", "SchNet"]
