"""Kernel-only training step: model fwd + bwd driven by fixed cotangents, R101 bs 8 480x640.
Usage: python tools/train_bench.py [preset] [B] [steps] [prec] [--profile]"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import build_ours, perturb_  # noqa: E402


def main():
    a = [v for v in sys.argv[1:] if not v.startswith("--")]
    preset = a[0] if len(a) > 0 else "PlaneRecNet_101_config"
    B = int(a[1]) if len(a) > 1 else 8
    steps = int(a[2]) if len(a) > 2 else 5
    prec = a[3] if len(a) > 3 else "bf16"
    prof = "--profile" in sys.argv
    torch.manual_seed(0)
    net = build_ours(preset)
    perturb_(net)
    net.train().cuda()
    from planerecnet_b200.train_engine import TrainEngine
    eng = net._train_engine = TrainEngine(prec)
    x = torch.randn(B, 3, 480, 640, device="cuda")
    cots = None
    times = []
    if "--graph" in sys.argv:
        from planerecnet_b200.train_engine import GraphedStep
        t0 = time.time()
        step = GraphedStep(eng, net, x, pack_in_graph="--no-pack" not in sys.argv)
        torch.cuda.synchronize()
        cap_s = time.time() - t0
        g = torch.Generator(device="cuda").manual_seed(1)
        mk = lambda t: torch.randn(t.shape, device="cuda", generator=g) / t[0].numel() ** 0.5  # noqa: E731
        m, cs, ks, d = step.outs
        cots = (mk(m), [mk(c) for c in cs], [mk(k) for k in ks], mk(d))
        for it in range(steps + 2):
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            step.forward(x)
            e1.record()
            step.backward(*cots)
            e2.record()
            torch.cuda.synchronize()
            if it >= 2:
                times.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        fwd = sum(t[0] for t in times) / len(times)
        bwd = sum(t[1] for t in times) / len(times)
        print("TRAIN_BENCH " + json.dumps({"preset": preset, "batch": B, "dtype": prec, "graph": True,
              "pack_in_graph": "--no-pack" not in sys.argv, "fwd_ms": fwd, "bwd_ms": bwd, "step_ms": fwd + bwd,
              "img_per_s": B * 1000.0 / (fwd + bwd), "launches": step.fwd_launches + step.bwd_launches, "capture_s": cap_s,
              "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
        return
    warm = int(os.environ.get('PRN_WARMUP', '2'))
    for it in range(steps + warm):
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        if prof and it == steps + warm - 1:
            eng.profile = []
        l0 = eng.launches
        e0.record()
        mask, cates, kerns, depth = eng.forward_train(net, x)
        e1.record()
        if cots is None:
            g = torch.Generator(device="cuda").manual_seed(1)
            mk = lambda t: torch.randn(t.shape, device="cuda", generator=g) / t[0].numel() ** 0.5  # noqa: E731
            cots = (mk(mask), [mk(c) for c in cates], [mk(k) for k in kerns], mk(depth))
        eng.seed_output_grads(*cots)
        grads = eng.backward()
        e2.record()
        torch.cuda.synchronize()
        if it >= warm:
            times.append((e0.elapsed_time(e1), e1.elapsed_time(e2)))
        nl = eng.launches - l0
        del grads
    fwd = sum(t[0] for t in times) / len(times)
    bwd = sum(t[1] for t in times) / len(times)
    out = {"preset": preset, "batch": B, "dtype": prec, "fwd_ms": fwd, "bwd_ms": bwd, "step_ms": fwd + bwd,
           "img_per_s": B * 1000.0 / (fwd + bwd), "launches_counted": nl,
           "mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    if prof and eng.profile:
        agg = {}
        for name, flops, s, e in eng.profile:
            d = agg.setdefault(name, [0.0, 0.0, 0])
            d[0] += s.elapsed_time(e)
            d[1] += flops
            d[2] += 1
        out["kinds"] = {k: {"ms": round(v[0], 3), "tflops": round(v[1] / v[0] / 1e9, 1) if v[0] > 0 else 0, "n": v[2]}
                        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}
    print("TRAIN_BENCH " + json.dumps(out))


if __name__ == "__main__":
    main()
