"""Configuration presets of the hot path — the same attribute names and values as the reference's
`data/config.py` (Config class :42-81, backbone presets :208-250, fpn_base :254-269, depth_fpn :273-282,
solov2_light :345-403, PlaneRecNet_*_config :407-528, set_cfg :533-540), restricted to what the dense
forward reads.  When the reference's own `data.config` is importable (drop-in use)
its `cfg` object can be passed to `PlaneRecNet(cfg)` instead: only attribute access is assumed."""
import copy

from .models.backbone import ResNetBackbone


# data/config.py:33-34: per-channel mean / std of the input images in BGR order (FastBaseTransform, data/augmentations.py:510-511)
MEANS = (103.94, 116.78, 123.68)
STD = (57.38, 57.12, 58.40)


class Config(object):
    """Attribute bag with copy()/replace(), mirroring data/config.py:42-81."""

    def __init__(self, config_dict):
        for k, v in config_dict.items():
            setattr(self, k, v)

    def copy(self, new_config_dict=None):
        ret = Config(vars(self))
        for k, v in (new_config_dict or {}).items():
            setattr(ret, k, v)
        return ret

    def replace(self, new_config_dict):
        if isinstance(new_config_dict, Config):
            new_config_dict = vars(new_config_dict)
        for k, v in new_config_dict.items():
            setattr(self, k, v)

    def print(self):
        for k, v in vars(self).items():
            print(k, " = ", v)


def _backbone(name, path, layers, dcn_layers=None, dcn_interval=None):
    args = (layers,) if dcn_layers is None else ((layers, dcn_layers) if dcn_interval is None else (layers, dcn_layers, dcn_interval))
    return Config({"name": name, "path": path, "type": ResNetBackbone, "args": args,
                   "selected_layers": list(range(2, 4))})


resnet101_dcn_inter3_backbone = _backbone("ResNet101_DCN_Interval3", "resnet101_reducedfc.pth",
                                          [3, 4, 23, 3], [0, 4, 23, 3], 3)
resnet50_dcnv2_backbone = _backbone("ResNet50_DCNv2", "resnet50-19c8e357.pth", [3, 4, 6, 3], [0, 4, 6, 3])

fpn_base = Config({"selected_layers": list(range(0, 4)), "start_level": 0, "num_features": 256,
                   "interpolation_mode": "bilinear", "high_level_mode": None, "relu_pred_layers": True})

depth_fpn = Config({"selected_layers": list(range(0, 4)), "skip_layers": list(range(0, 4)), "use_refle": True})

solov2_light = Config({
    "num_kernels": 128, "masks_in_features": ["p2", "p3", "p4", "p5"], "masks_channels": 128,
    "num_masks": 128, "instance_in_features": ["p2", "p3", "p4", "p5"], "instance_channels": 256,
    "fpn_instance_strides": [8, 8, 16, 32],
    "fpn_scale_ranges": ((1, 128), (64, 256), (128, 512), (256, 2048)),
    "num_grids": [40, 36, 24, 16], "num_instance_convs": 3, "use_dcn_in_instance": False, "sigma": 0.2,
    "nms_pre": 500, "score_thr": 0.1, "nms_type": "matrix", "mask_thr": 0.1, "update_thr": 0.15,
    "nms_kernel": "gaussian", "nms_sigma": 2, "top_k": 100,
    "use_coord_conv": True, "norm": "GN", "focal_loss_init_pi": 0.01,
})

PlaneRecNet_101_config = Config({
    "name": "PlaneRecNet_101",
    "num_classes": 2,
    "freeze_bn": False,
    "backbone": resnet101_dcn_inter3_backbone,
    "fpn": fpn_base,
    "depth": depth_fpn,
    "solov2": solov2_light,
    "max_size": 640,
    "device": "cuda",
    # loss weights (data/config.py:459-468, 511-514) — carried for callers, unused by the dense forward
    "dice_weight": 3.0, "focal_weight": 1.0, "depth_weight": 5.0, "use_lava_loss": True, "use_plane_loss": True,
    "lava_weight": 1.0, "pln_weight": 1.0, "focal_gamma": 2.0, "focal_alpha": 0.25,
    # scannet_dataset (data/config.py:113-136): only what the loss reads
    "dataset": Config({"name": "ScanNetDataset", "depth_resolution": 1 / 1000, "min_depth": 1 / 1000, "max_depth": 40}),
})

PlaneRecNet_50_config = PlaneRecNet_101_config.copy({"name": "PlaneRecNet_50", "backbone": resnet50_dcnv2_backbone})

_PRESETS = {"PlaneRecNet_101_config": PlaneRecNet_101_config, "PlaneRecNet_50_config": PlaneRecNet_50_config}

# the active config, mutated in place like the reference's global `cfg`
cfg = PlaneRecNet_101_config.copy()


def set_cfg(config_name: str):
    """Sets the active config in place (data/config.py:533-540)."""
    if config_name not in _PRESETS:
        raise KeyError(f"unknown config {config_name!r}; available: {sorted(_PRESETS)}")
    src = _PRESETS[config_name]
    cfg.replace({k: copy.copy(v) for k, v in vars(src).items()})
    cfg.solov2 = src.solov2.copy()
    if cfg.name is None:
        cfg.name = config_name.split("_config")[0]
    return cfg
