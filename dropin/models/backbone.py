from planerecnet_b200.models.backbone import Bottleneck, ResNetBackbone, construct_backbone  # noqa: F401
