"""Drop-in for the reference's `models.functions.losses` (losses.py:12-198): same class name, constructor and forward
signature; the dense parts run on libprn_b200 (see planerecnet_b200/losses.py)."""
from planerecnet_b200.losses import PlaneRecNetLoss as _PlaneRecNetLoss


class PlaneRecNetLoss(_PlaneRecNetLoss):
    def __init__(self):
        from data.config import cfg      # the reference's live config object (train.py:16, losses.py:5)
        super().__init__(cfg)
