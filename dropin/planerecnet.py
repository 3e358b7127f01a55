"""Drop-in for the reference's top-level `planerecnet` module."""
from planerecnet_b200.planerecnet import DepthDecoder_FPN, PlaneRecNet, SOLOv2InsHead, SOLOv2MaskHead  # noqa: F401
