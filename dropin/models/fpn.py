from planerecnet_b200.models.fpn import FPN  # noqa: F401
