"""Times the UNMODIFIED reference (baseline/_ref, staged by baseline/stage_reference.py) through its own public API
(`planerecnet.PlaneRecNet`, `models.functions.losses.PlaneRecNetLoss`) on the CPU or on the GPU and prints ONE JSON line.
Runs in its own process (bench.py spawns it) so that the reference's top-level module names (`planerecnet`, `models`,
`data`, `utils`) never mix with the repo's package and none of the repo's kernels / .so files are loaded here.

  python baseline/ref_runner.py --device cpu|cuda --mode fwd_dense|fwd_e2e|train_cot|train_loss
         [--preset PlaneRecNet_101_config] [--batch 8] [--steps 5] [--warmup 2] [--amp] [--channels-last] [--no-tf32]

Modes:
  fwd_dense   eval-mode BatchNorm, dense forward only (backbone -> FPN -> heads -> depth decoder): the top-level module's
              `training` flag alone is set so that planerecnet.py:99-103 returns the dense tuple (no reference file is edited)
  fwd_e2e     net.eval()(x): dense forward + inference bookkeeping (planerecnet.py:104-111)
  train_cot   net.train() forward + backward driven by fixed seeded cotangents of the 10 outputs (config 4 (i))
  train_loss  net.train() forward + PlaneRecNetLoss + backward (config 4 (ii)); on the CPU `.cuda()` is patched to identity
              (SURVEY §8c runtime patches)
Only runtime patches, exactly those of SURVEY §8c; the reference files are byte-identical to /root/reference (manifest check)."""
import argparse
import importlib.util
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.path.join(HERE, "_ref")


def load_synth():
    spec = importlib.util.spec_from_file_location("_prn_synth", os.path.join(ROOT, "planerecnet_b200", "utils", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--mode", default="fwd_dense", choices=["fwd_dense", "fwd_e2e", "train_cot", "train_loss"])
    ap.add_argument("--preset", default="PlaneRecNet_101_config")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--amp", action="store_true", help="torch.autocast(float16) around the forward (GPU only)")
    ap.add_argument("--channels-last", action="store_true")
    ap.add_argument("--no-tf32", action="store_true")
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()

    sys.path.insert(0, HERE)
    import stage_reference as SR
    if not os.path.isdir(REF):
        print(json.dumps({"unavailable": "baseline/_ref is not staged (python baseline/stage_reference.py needs /root/reference)"}))
        return 0
    manifest_ok = SR.verify()

    import numpy as np
    import torch
    cuda = a.device == "cuda"
    if cuda and not torch.cuda.is_available():
        print(json.dumps({"unavailable": "no CUDA device"}))
        return 0
    cores = os.cpu_count() or 1
    if not cuda:
        torch.set_num_threads(a.threads or cores)
        torch.cuda.current_device = lambda: 0                      # planerecnet.py:18 touches CUDA at import
        torch.Tensor.cuda = lambda self, *args, **kw: self          # hard-coded .cuda() in vnl.py / losses.py
    else:
        torch.backends.cudnn.benchmark = True
        torch.backends.cuda.matmul.allow_tf32 = not a.no_tf32
        torch.backends.cudnn.allow_tf32 = not a.no_tf32
    sys.path.insert(0, REF)
    S = load_synth()
    import planerecnet as ref_mod                                   # the reference module, unmodified
    from data.config import cfg, set_cfg
    from utils import timer
    timer.disable_all()
    set_cfg(a.preset)
    torch.manual_seed(0)
    net = S.perturb_(ref_mod.PlaneRecNet(cfg))
    x = S.make_input(a.batch, a.height, a.width, 0)
    dev = torch.device(a.device)
    net = net.to(dev)
    x = x.to(dev)
    if a.channels_last:
        net = net.to(memory_format=torch.channels_last)
        x = x.contiguous(memory_format=torch.channels_last)

    def sync():
        if cuda:
            torch.cuda.synchronize()

    def fwd(xx):
        if cuda and a.amp:
            with torch.autocast("cuda", dtype=torch.float16):
                return net(xx)
        return net(xx)

    crit = gts = gt_depth = cots = None
    if a.mode == "fwd_dense":
        net.eval()
        net.training = True            # top-level flag only: BatchNorm modules stay in eval mode, forward returns the dense tuple
    elif a.mode == "fwd_e2e":
        net.eval()
    else:
        net.train()
        if a.mode == "train_loss":
            if cuda:
                try:
                    torch.set_default_tensor_type("torch.cuda.FloatTensor")     # train.py:121-122
                except Exception:
                    torch.set_default_device("cuda")
            from models.functions.losses import PlaneRecNetLoss
            crit = PlaneRecNetLoss()
            gts, gt_depth = S.make_gt(a.batch, a.height, a.width, seed=0)
            gts = [{k: v.to(dev) for k, v in g.items()} for g in gts]
            gt_depth = gt_depth.to(dev)

    def step():
        if a.mode in ("fwd_dense", "fwd_e2e"):
            with torch.no_grad():
                return fwd(x)
        for p in net.parameters():
            p.grad = None
        outs = fwd(x)
        if a.mode == "train_cot":
            nonlocal cots
            if cots is None:
                cots = S.make_cotangents(outs, seed=1, device=dev)
            m, cs, ks, d = outs
            torch.autograd.backward([m] + list(cs) + list(ks) + [d], [cots[0]] + cots[1] + cots[2] + [cots[3]])
            return None
        np.random.seed(0)
        losses = crit(net, outs[0], outs[1], outs[2], outs[3], gts, gt_depth)
        losses = {k: v.mean() for k, v in losses.items()}         # train.py:347-348
        loss = sum(losses[k] for k in losses)
        loss.backward()
        return {k: float(v) for k, v in losses.items()}

    last = None
    for _ in range(a.warmup):
        last = step()
    sync()
    t0 = time.perf_counter()
    if cuda:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    for _ in range(a.steps):
        last = step()
    if cuda:
        e1.record()
        sync()
        ms = e0.elapsed_time(e1) / a.steps
    else:
        ms = (time.perf_counter() - t0) / a.steps * 1e3
    out = {"impl": "reference", "device": a.device, "mode": a.mode, "preset": a.preset, "batch": a.batch, "steps": a.steps,
           "warmup": a.warmup, "ms_per_step": ms, "images_per_s": a.batch / (ms / 1e3), "manifest_ok": manifest_ok,
           "cores": (a.threads or cores) if not cuda else None,
           "dtype": ("fp16 autocast" if a.amp else ("fp32 (cuDNN/cuBLAS TF32 allowed)" if cuda and not a.no_tf32 else "fp32")),
           "channels_last": bool(a.channels_last), "torch": torch.__version__}
    if isinstance(last, dict):
        out["losses"] = last
    if cuda:
        out["gpu"] = torch.cuda.get_device_name(0)
        out["max_mem_gb"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
