from planerecnet_b200.models.dcn import DeformableConv2d  # noqa: F401
