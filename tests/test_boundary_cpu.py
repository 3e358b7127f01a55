"""CPU tests of the drop-in boundary: module API / state_dict contract, config presets,
and that libprn_b200.so loads and exports every symbol include/prn_b200.h declares.  No compute call
is made here (there is no GPU in the build container)."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

import helpers as H

ROOT = H.ROOT


def test_header_symbols_exported_by_library():
    from planerecnet_b200 import _lib
    from planerecnet_b200.csrc import build as B
    lib_path = B.build()
    hdr = open(os.path.join(ROOT, "include", "prn_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(prn_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(_lib.EXPORTS), "include/prn_b200.h and _lib.EXPORTS disagree"
    nm = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (prn_[a-z0-9_]+)", nm))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"libprn_b200.so does not export {missing}"
    l = ctypes.CDLL(lib_path)           # loads without a GPU
    l.prn_abi_version.restype = ctypes.c_int
    assert l.prn_abi_version() == 1


def test_conv_descriptor_struct_matches_header_layout():
    """ctypes mirror of PrnConv must have the same field order as the C struct."""
    from planerecnet_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "prn_b200.h")).read()
    body = hdr[hdr.index("typedef struct PrnConv {") + len("typedef struct PrnConv {"):hdr.index("} PrnConv;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        m = re.match(r"(?:const\s+)?(?:void|float|int32_t)\s*\*?\s*(.+)$", decl)
        if m:
            names += [n.strip().lstrip("*") for n in m.group(1).split(",")]
    assert names == [f[0] for f in _lib.PrnConv._fields_]


def test_wgrad_descriptor_struct_matches_header_layout():
    from planerecnet_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "prn_b200.h")).read()
    body = hdr[hdr.index("typedef struct PrnWgrad {") + len("typedef struct PrnWgrad {"):hdr.index("} PrnWgrad;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        m = re.match(r"(?:const\s+)?(?:void|float|int32_t)\s*\*?\s*(.+)$", decl.strip())
        if m:
            names += [n.strip().lstrip("*") for n in m.group(1).split(",")]
    assert names == [f[0] for f in _lib.PrnWgrad._fields_]
    d = _lib.PrnWgrad()
    out8 = (ctypes.c_int32 * 8)()
    assert _lib.lib().prn_conv2d_wgrad_plan(ctypes.byref(d), out8) == -1      # rejected on the host, no GPU needed


def test_invalid_descriptor_is_rejected_without_gpu():
    from planerecnet_b200 import _lib
    l = _lib.lib()
    d = _lib.PrnConv()
    rc = l.prn_conv2d_plan(ctypes.byref(d), None, None, None)
    assert rc == -1 and b"conv" in l.prn_last_error()
    with pytest.raises(_lib.PrnError):
        _lib.check(rc, "prn_conv2d_plan")


@pytest.mark.parametrize("preset,n_keys,n_params", [("PlaneRecNet_50_config", 520, 38.55e6),
                                                    ("PlaneRecNet_101_config", 816, 57.36e6)])
def test_state_dict_contract(preset, n_keys, n_params):
    net = H.build_ours(preset)
    sd = net.state_dict()
    assert len(sd) == n_keys
    total = sum(p.numel() for p in net.parameters())
    assert abs(total - n_params) < 0.01e6
    keys = list(sd)
    assert keys[0].startswith("backbone.layers.0.0.conv1")            # ModuleList registered before conv1
    assert keys.index("backbone.conv1.weight") > keys.index("backbone.layers.3.2.bn3.num_batches_tracked")
    for k in ["fpn.lateral_convs.3.bias", "fpn.fpn_convs.0.weight", "inst_head.kernel_tower.0.weight",
              "inst_head.cate_tower.7.bias", "inst_head.cate_pred.bias", "inst_head.kernel_pred.weight",
              "mask_head.convs_all_levels.3.conv2.1.bias", "mask_head.conv_pred.0.weight",
              "depth_decoder.latlayer1.weight", "depth_decoder.conv1.1.weight", "depth_decoder.conv1.2.running_var",
              "depth_decoder.deconv4.2.weight", "depth_decoder.deconv4.3.num_batches_tracked",
              "depth_decoder.depth_pred.1.bias", "depth_decoder.conv1x1.0.weight", "depth_decoder.refine_conv.1.weight"]:
        assert k in sd, k
    assert tuple(sd["inst_head.kernel_tower.0.weight"].shape) == (256, 258, 3, 3)
    assert tuple(sd["mask_head.convs_all_levels.3.conv0.0.weight"].shape) == (128, 258, 3, 3)
    assert tuple(sd["depth_decoder.conv1x1.0.weight"].shape) == (256, 3728, 1, 1)
    # the five top-level sub-modules partition the parameters (Adam groups of train.py:251-256)
    parts = sum(sum(p.numel() for p in m.parameters()) for m in
                (net.backbone, net.fpn, net.inst_head, net.mask_head, net.depth_decoder))
    assert parts == total


def test_dcn_placement_rule():
    """models/backbone.py:170,184 with the presets' arguments."""
    from planerecnet_b200.models.dcn import DeformableConv2d

    def dcn_blocks(preset):
        net = H.build_ours(preset)
        return [(s, b) for s, layer in enumerate(net.backbone.layers) for b, blk in enumerate(layer)
                if isinstance(blk.conv2, DeformableConv2d)]

    r101 = dcn_blocks("PlaneRecNet_101_config")
    assert r101 == [(1, 0), (1, 3)] + [(2, b) for b in range(0, 23, 3)] + [(3, 0)]
    r50 = dcn_blocks("PlaneRecNet_50_config")
    assert len(r50) == 13 and all(s >= 1 for s, _ in r50)
    net = H.build_ours("PlaneRecNet_50_config")
    blk = net.backbone.layers[1][0].conv2
    assert blk.offset_conv.weight.abs().sum() == 0 and blk.modulator_conv.bias.abs().sum() == 0
    assert blk.regular_conv.bias is not None and blk.stride == 2


def test_config_presets_and_set_cfg():
    from planerecnet_b200.config import cfg, set_cfg
    set_cfg("PlaneRecNet_101_config")
    assert cfg.name == "PlaneRecNet_101" and cfg.backbone.args == ([3, 4, 23, 3], [0, 4, 23, 3], 3)
    assert cfg.solov2.num_grids == [40, 36, 24, 16] and cfg.solov2.fpn_instance_strides == [8, 8, 16, 32]
    set_cfg("PlaneRecNet_50_config")
    assert cfg.name == "PlaneRecNet_50" and cfg.backbone.args == ([3, 4, 6, 3], [0, 4, 6, 3])
    cfg.solov2.replace({"score_thr": 0.3})          # callers override before construction
    assert cfg.solov2.score_thr == 0.3
    set_cfg("PlaneRecNet_50_config")
    assert cfg.solov2.score_thr == 0.1
    with pytest.raises(KeyError):
        set_cfg("nope")


def test_forward_without_gpu_fails_loudly():
    """No CPU fallback: a CPU tensor (or a missing CUDA library) must raise, not silently compute."""
    net = H.build_ours("PlaneRecNet_50_config").eval()
    with pytest.raises(Exception) as ei:
        net(torch.zeros(1, 3, 64, 64))
    assert "CUDA" in str(ei.value) or "cuda" in str(ei.value)


def test_training_mode_has_no_cpu_path_either():
    """net.train() routes to the sm_100a training executor; like the eval path it fails loudly on a CPU tensor."""
    from planerecnet_b200 import _lib
    net = H.build_ours("PlaneRecNet_50_config").train()
    with pytest.raises((_lib.PrnError, RuntimeError, AssertionError)):
        net(torch.zeros(1, 3, 64, 64))


def test_dgrad_weight_packing_layout():
    """Input-gradient weights: rows = input channels, K = (ky, kx, cout padded) with the taps flipped."""
    from planerecnet_b200 import ops, _lib
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)          # [Cout=2, Cin=3, 3, 3]
    p = ops.pack_dgrad_weight(w, cout_pad=64, n_pad=16, dtype=_lib.PRN_F16)
    assert p.shape == (16, 9 * 64)
    # row c, k = (ky*3 + kx)*64 + o holds w[o, c, 2-ky, 2-kx]
    assert float(p[1, (0 * 3 + 2) * 64 + 1]) == float(w[1, 1, 2, 0])
    assert p[3:].abs().sum() == 0 and float(p[0, 2]) == 0
    g = torch.arange(4 * 9 * 64, dtype=torch.float32).reshape(4, 9 * 64)
    u = ops.unpack_wgrad(g, (2, 3, 3, 3), [(3, 64)])
    assert u.shape == (2, 3, 3, 3) and float(u[1, 2, 1, 0]) == float(g[1, (1 * 3 + 0) * 64 + 2])


def test_weight_packing_layout():
    from planerecnet_b200 import ops, _lib
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    p = ops.pack_conv_weight(w, [(3, 64)], 16, _lib.PRN_F16)
    assert p.shape == (16, 9 * 64) and p.dtype == torch.float16
    # k = (ky*3 + kx)*64 + c
    assert float(p[1, (1 * 3 + 2) * 64 + 2]) == float(w[1, 2, 1, 2])
    assert float(p[0, 3]) == 0 and p[2:].abs().sum() == 0
    q = ops.pack_conv_weight(torch.ones(4, 6, 1, 1), [(2, 64), (4, 64)], 16, _lib.PRN_BF16, scale=torch.tensor([1., 2., 3., 4.]))
    assert q.shape == (16, 128) and float(q[2, 0]) == 3 and float(q[2, 2]) == 0 and float(q[2, 64 + 3]) == 3


def test_init_weights_semantics(tmp_path):
    """planerecnet.py:130-145: backbone from checkpoint (strict=False, key rename), xavier elsewhere,
    cate_pred bias = -log(99)."""
    net = H.build_ours("PlaneRecNet_50_config")
    ck = {"conv1.weight": torch.full((64, 3, 7, 7), 0.5), "layer1.0.conv1.weight": torch.full((64, 64, 1, 1), 0.25),
          "fc.weight": torch.zeros(10, 10)}
    path = tmp_path / "bb.pth"
    torch.save(ck, path)
    net.init_weights(str(path))
    assert float(net.backbone.conv1.weight[0, 0, 0, 0]) == 0.5
    assert float(net.backbone.layers[0][0].conv1.weight[0, 0, 0, 0]) == 0.25
    assert abs(float(net.inst_head.cate_pred.bias[0]) + 4.59512) < 1e-4
    assert float(net.fpn.lateral_convs[0].bias.abs().sum()) == 0
    net.save_weights(str(tmp_path / "w.pth"))
    net2 = H.build_ours("PlaneRecNet_50_config", seed=5)
    net2.load_weights(str(tmp_path / "w.pth"))
    assert torch.equal(net2.fpn.fpn_convs[2].weight, net.fpn.fpn_convs[2].weight)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference checkout only exists in the build container")
def test_dropin_shims_work_with_the_reference_config():
    """with dropin/ ahead of the reference on sys.path, the reference's own data/config.py builds
    our model (construction only; no GPU here)."""
    code = (
        "import sys; sys.path[:0] = [%r, %r, '/root/reference']\n"
        "import torch\n"
        "from data.config import cfg, set_cfg\n"
        "import models.backbone as mb\n"
        "assert mb.__name__ == 'models.backbone' and 'planerecnet_b200' in mb.ResNetBackbone.__module__\n"
        "set_cfg('PlaneRecNet_50_config')\n"
        "from planerecnet import PlaneRecNet\n"
        "net = PlaneRecNet(cfg)\n"
        "assert len(net.state_dict()) == 520\n"
        "from models.functions.nms import matrix_nms, point_nms\n"
        "import models.functions.funcs as f; assert abs(f.bias_init_with_prob(0.01) + 4.59512) < 1e-4\n"
        "print('ok')\n" % (os.path.join(ROOT, "dropin"), ROOT))
    out = subprocess.run([os.sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_fused_adam_tables_on_cpu():
    """Host logic of planerecnet_b200.optim.FusedAdam (no launch): pointer table, strides of 1-D gradient views, chunking,
    per-group learning rates and their in-place refresh (train.py:251-256, 415-420)."""
    from planerecnet_b200.optim import FusedAdam, _CHUNK
    a, b, c = (torch.nn.Parameter(torch.zeros(s)) for s in ((3, 4), (_CHUNK + 5,), (7,)))
    frozen = torch.nn.Parameter(torch.zeros(2), requires_grad=False)
    opt = FusedAdam([{"params": [a, frozen], "lr": 5e-4}, {"params": [b, c]}], lr=1e-4)
    acc = torch.zeros(7, 2)
    grads = {id(a): torch.ones(3, 4), id(b): torch.ones(_CHUNK + 5), id(c): acc[:, 1], id(frozen): torch.ones(2)}
    opt.prepare(grads)
    table, numel, lrs, chunks, n_chunks, _ = opt._tabs
    assert table.view(-1, 5).shape[0] == 3                      # the frozen parameter is skipped
    assert table.view(-1, 5)[:, 4].tolist() == [1, 1, 2]         # column view of a [C,2] accumulator: element stride 2
    assert table.view(-1, 5)[2, 1].item() == acc[:, 1].data_ptr()
    assert numel.tolist() == [12, _CHUNK + 5, 7] and n_chunks == 4
    assert chunks.tolist() == [[0, 0], [1, 0], [1, 1], [2, 0]]
    assert torch.allclose(lrs, torch.tensor([5e-4, 1e-4, 1e-4]))
    lr_tensor = opt._tabs[2]
    opt.param_groups[1]["lr"] = 2e-5                             # set_lr: same table object, refreshed in place
    opt.prepare(grads)
    assert opt._tabs[2] is lr_tensor and torch.allclose(lr_tensor, torch.tensor([5e-4, 2e-5, 2e-5]))
    assert set(opt.state[id(a)]) == {"exp_avg", "exp_avg_sq"}


def test_subpixel_weights_reproduce_upsample_reflect_conv():
    """ops.subpixel_weights: Upsample(x2, nearest) -> ReflectionPad2d(1) -> Conv2d(3x3) (planerecnet.py:540-567) equals a 3x3
    convolution with replicate padding at the low resolution followed by a pixel shuffle of the four phase blocks."""
    import torch
    import torch.nn.functional as F
    from planerecnet_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 5, 7, 9, generator=g, dtype=torch.float64)
    w = torch.randn(6, 5, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.pad(F.interpolate(x, scale_factor=2, mode="nearest"), (1, 1, 1, 1), mode="reflect"), w)
    wc = ops.subpixel_weights(w.float()).double()
    # subpixel_weights sums in fp32: compare against the same sums in fp64 to 1e-6
    low = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), wc).reshape(2, 2, 2, 6, 7, 9)       # [B, a, b, C, H, W]
    got = low.permute(0, 3, 4, 1, 5, 2).reshape(2, 6, 14, 18)
    assert float((got - ref).abs().max()) < 1e-5 * float(ref.abs().max())
