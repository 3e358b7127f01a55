"""Run the reference's OWN scripts against this repo's modules — the "drops into train.py /
simple_inference.py unchanged" claim, on hardware.

  python tools/run_reference_script.py simple_inference [--config PlaneRecNet_50_config]
  python tools/run_reference_script.py train_loop [--config PlaneRecNet_101_config] [--iters 3] [--batch 2]

* sys.path = [<repo>/dropin, <repo>, <reference>]: `planerecnet`, `models.backbone|dcn|fpn`, `models.functions.nms|funcs|losses`
  resolve to this repo; `data/`, `utils/`, `models/functions/vnl.py` and the scripts themselves come from the unmodified
  reference checkout (baseline/_ref, staged by baseline/stage_reference.py; /root/reference in the build container).
* third-party names the reference imports but this image lacks get tiny `sys.modules` shims (no reference file is edited):
  `tensorboardX` (train.py:13), `pycocotools` (data/datasets.py:10), `numpy.core.numeric.NaN` (simple_inference.py:10,
  removed in numpy 2), `np.float` (data/datasets.py:105).
* `simple_inference`: reference `simple_inference.py` __main__ (lines 327-369) on data/example_nyu.jpg with a random-init
  checkpoint written by `net.save_weights` (the published checkpoints are not available offline).
* `train_loop`: the reference `train.py` module is executed up to its definitions (argument parsing, cfg, NetLoss,
  CustomDataParallel: lines 1-226), then the body of its training loop (lines 251-256, 264-265, 344-354: Adam over five
  parameter groups, CustomDataParallel(NetLoss(net, PlaneRecNetLoss())), zero_grad / forward+loss / backward / step) runs for
  `--iters` iterations on synthetic ScanNet-shaped batches (its dataset classes need the ScanNet files).
Prints one JSON line per run."""
import argparse
import json
import os
import runpy
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    for cand in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.isfile(os.path.join(cand, "train.py")):
            return cand
    raise SystemExit("reference checkout not found: run `python baseline/stage_reference.py` in the build container")


def install_shims():
    import numpy as np
    tbx = types.ModuleType("tensorboardX")

    class SummaryWriter:                      # train.py:13, 271-276: only constructed and fed scalars
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass

        def close(self):
            pass

    tbx.SummaryWriter = SummaryWriter
    sys.modules.setdefault("tensorboardX", tbx)
    coco = types.ModuleType("pycocotools")
    coco_mask = types.ModuleType("pycocotools.mask")      # data/datasets.py:10: annotation decoding of the real datasets only
    coco.mask = coco_mask
    sys.modules.setdefault("pycocotools", coco)
    sys.modules.setdefault("pycocotools.mask", coco_mask)
    import numpy.core.numeric as ncn          # simple_inference.py:10
    if not hasattr(ncn, "NaN"):
        ncn.NaN = float("nan")
    if not hasattr(np, "float"):
        np.float = float                      # data/datasets.py:105


def setup_paths(ref):
    for p in (ref, ROOT, os.path.join(ROOT, "dropin")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)


def run_simple_inference(a):
    import torch
    ref = find_reference()
    setup_paths(ref)
    install_shims()
    from data.config import cfg, set_cfg          # reference's config
    import planerecnet                              # ours, through dropin/
    assert planerecnet.PlaneRecNet.__module__.startswith("planerecnet_b200"), planerecnet.PlaneRecNet.__module__
    set_cfg(a.config)
    torch.manual_seed(0)
    tmp = tempfile.mkdtemp(prefix="prn_dropin_")
    ckpt = os.path.join(tmp, "PlaneRecNet_random_init.pth")
    net = planerecnet.PlaneRecNet(cfg)
    keys = list(net.state_dict().keys())
    with torch.no_grad():
        # the script looks class names up in cfg.dataset.class_names = ('plane',) (simple_inference.py:106): a random-init model
        # must not emit class 1, which no trained checkpoint does
        net.inst_head.cate_pred.bias[1] = -20.0
    net.save_weights(ckpt)
    del net
    out_png = os.path.join(tmp, "out.png")
    img = os.path.join(ref, "data", "example_nyu.jpg")
    sys.argv = ["simple_inference.py", f"--config={a.config}", f"--trained_model={ckpt}", f"--image={img}:{out_png}",
                "--score_threshold=0.15"]
    cwd = os.getcwd()
    os.chdir(tmp)
    try:
        runpy.run_path(os.path.join(ref, "simple_inference.py"), run_name="__main__")
    finally:
        os.chdir(cwd)
    produced = sorted(os.listdir(tmp))
    ok = os.path.exists(out_png) and os.path.getsize(out_png) > 0
    reloaded = torch.load(ckpt, map_location="cpu")
    print(json.dumps({"script": "simple_inference.py", "config": a.config, "ok": bool(ok), "outputs": produced,
                      "state_dict_keys": len(keys), "keys_roundtrip": list(reloaded.keys()) == keys,
                      "model_class": planerecnet.PlaneRecNet.__module__}))
    return 0 if ok else 1


def run_train_loop(a):
    import torch
    ref = find_reference()
    setup_paths(ref)
    install_shims()
    sys.argv = ["train.py", f"--config={a.config}", f"--batch_size={a.batch}", "--no_autoscale"]
    g = runpy.run_path(os.path.join(ref, "train.py"), run_name="reference_train")      # definitions only: no __main__ body
    cfg, args = g["cfg"], g["args"]
    PlaneRecNet, PlaneRecNetLoss = g["PlaneRecNet"], g["PlaneRecNetLoss"]
    assert PlaneRecNet.__module__.startswith("planerecnet_b200") and PlaneRecNetLoss.__mro__[1].__module__.startswith("planerecnet_b200")
    import numpy as np
    from planerecnet_b200.utils.synth import make_gt, make_input
    torch.manual_seed(0)
    prn_net = PlaneRecNet(cfg)
    net = prn_net
    net.train()
    optimizer = torch.optim.Adam([                                   # train.py:251-256
        {"params": net.backbone.parameters(), "lr": 5 * args.lr}, {"params": net.fpn.parameters(), "lr": args.lr},
        {"params": net.inst_head.parameters(), "lr": args.lr}, {"params": net.mask_head.parameters(), "lr": args.lr},
        {"params": net.depth_decoder.parameters(), "lr": 2 * args.lr}], lr=args.lr)
    criterion = PlaneRecNetLoss()                                    # train.py:258
    net = g["CustomDataParallel"](g["NetLoss"](net, criterion))      # train.py:264
    net = net.cuda()
    before = {k: v.detach().clone() for k, v in prn_net.state_dict().items()}
    hist = []
    for it in range(a.iters):
        # a datum as detection_collate (data/datasets.py) hands it over: lists of per-image CPU tensors
        x = make_input(a.batch, 480, 640, seed=it)
        gts, gtd = make_gt(a.batch, 480, 640, seed=it)
        datum = ([x[i] for i in range(a.batch)], gts, [gtd[i] for i in range(a.batch)])
        np.random.seed(it)
        optimizer.zero_grad()                                        # train.py:344
        losses = net(datum)                                          # train.py:347
        losses = {k: (v).mean() for k, v in losses.items()}          # train.py:348
        loss = sum([losses[k] for k in losses])                      # train.py:349
        loss.backward()                                              # train.py:352
        if torch.isfinite(loss).item():                              # train.py:353-354
            optimizer.step()
        hist.append({k: float(v) for k, v in losses.items()})
    after = prn_net.state_dict()
    moved = sum(1 for k in before if before[k].is_floating_point() and not torch.equal(before[k], after[k]))
    finite = all(np.isfinite(v) for h in hist for k, v in h.items() if k != "pln") and all(np.isfinite(h["ins"]) for h in hist)
    tmp = tempfile.mkdtemp(prefix="prn_dropin_")
    path = os.path.join(tmp, "w.pth")
    prn_net.save_weights(path)                                       # train.py:386 (SavePath) -> planerecnet.py:121-123
    twin = PlaneRecNet(cfg)
    twin.load_weights(path)                                          # strict load of all keys
    same = all(torch.equal(v.cpu(), twin.state_dict()[k].cpu()) for k, v in after.items())
    print(json.dumps({"script": "train.py loop", "config": a.config, "iters": a.iters, "batch": a.batch, "losses": hist,
                      "finite": bool(finite), "tensors_updated": moved, "state_dict_keys": len(after), "save_load_roundtrip": bool(same)}))
    return 0 if (finite and moved > 0 and same) else 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["simple_inference", "train_loop"])
    ap.add_argument("--config", default=None)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--batch", type=int, default=2)
    a = ap.parse_args()
    if a.config is None:
        a.config = "PlaneRecNet_50_config" if a.what == "simple_inference" else "PlaneRecNet_101_config"
    return run_simple_inference(a) if a.what == "simple_inference" else run_train_loop(a)


if __name__ == "__main__":
    sys.exit(main())
