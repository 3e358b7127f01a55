"""Host timeline of the end-to-end training step bench.py times as `e2e` (pinned host batch -> H2D -> prepare -> net(x) -> loss ->
backward -> D2H of the loss terms): where the HOST spends the step (no synchronisation between the phases, so a phase that waits
for the device shows up where the wait happens), then a cProfile of three steps.   python tools/e2e_train_profile.py [numpy|device]"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200.config import cfg, set_cfg  # noqa: E402
from planerecnet_b200.losses import PlaneRecNetLoss  # noqa: E402
from planerecnet_b200.planerecnet import PlaneRecNet  # noqa: E402
from planerecnet_b200.utils.synth import make_gt, make_input, perturb_  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "numpy"
    B = 8
    set_cfg("PlaneRecNet_101_config")
    torch.manual_seed(0)
    net = perturb_(PlaneRecNet(cfg)).train().cuda()
    net.use_train_graph = True
    crit = PlaneRecNetLoss(cfg, vnl_sampling=mode)
    if os.environ.get("PRN_NO_PLANE") == "1":          # experiment: what the plane term costs the step
        crit.use_plane = False
    if os.environ.get("PRN_NO_LAVA") == "1":
        crit.use_lava = False
    x_host = make_input(B, 480, 640, 0).pin_memory()
    gts_h, gtd_h = make_gt(B, 480, 640, seed=0)
    gts_h = [{k: v.pin_memory() for k, v in g.items()} for g in gts_h]
    gtd_h = gtd_h.pin_memory()
    params = [p for p in net.parameters() if p.requires_grad]
    loss_host = torch.empty(5).pin_memory()
    marks = {}

    copy_stream = torch.cuda.Stream()

    def upload():
        with torch.cuda.stream(copy_stream):
            x = x_host.cuda(non_blocking=True)
            gts = [{k: v.cuda(non_blocking=True) for k, v in g.items()} for g in gts_h]
            gtd = gtd_h.cuda(non_blocking=True)
            crit.prepare(gts)
        return x, gts, gtd

    def loop(n, lookahead):
        """bench.py's e2e loop (user_loop) with host time stamps between the phases."""
        np.random.seed(0)
        nxt = upload() if lookahead else None
        for i in range(n):
            t = [time.perf_counter()]
            batch = nxt if lookahead else upload()
            main = torch.cuda.current_stream()
            main.wait_stream(copy_stream)
            x, gts, gtd = batch
            for p in params:
                p.grad = None
            t.append(time.perf_counter())
            outs = net(x)
            t.append(time.perf_counter())
            losses = crit(net, outs[0], outs[1], outs[2], outs[3], gts, gtd)
            losses = {k: v.mean() for k, v in losses.items()}
            t.append(time.perf_counter())
            nxt = upload() if (lookahead and i + 1 < n) else None
            t.append(time.perf_counter())
            sum(losses[k] for k in losses).backward()
            t.append(time.perf_counter())
            vals = torch.stack([losses[k].detach().float() for k in ("ins", "cat", "dpt", "pln", "lav")])
            loss_host.copy_(vals, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            t.append(time.perf_counter())
            for name, a, b in zip(("upload / wait", "net(x)", "loss", "prefetch next", "backward", "d2h + final sync"), t[:-1], t[1:]):
                marks[name] = marks.get(name, 0.0) + (b - a)
        return loss_host.tolist()

    def step(_prepare=True):
        return loop(1, False)

    for lookahead in (True, False):
        loop(3, lookahead)
        marks.clear()
        torch.cuda.synchronize()
        n = 6
        t0 = time.perf_counter()
        lv = loop(n, lookahead)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n * 1e3
        print(f"[{mode}, look-ahead {lookahead}] {dt:.1f} ms per step = {B / dt * 1e3:.1f} images/s | host: " +
              " | ".join(f"{k} {v / n * 1e3:.1f}" for k, v in marks.items()) + f" | losses {[round(v, 4) for v in lv]}")
    # device side of the same step: kernel time by name (CUPTI through torch.profiler), network graphs vs everything else
    import collections
    import json
    import tempfile
    from torch.profiler import ProfilerActivity, profile
    reps = 2
    reps = 3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        loop(reps, True)
        torch.cuda.synchronize()
    tmp = tempfile.mktemp(suffix=".json")
    prof.export_chrome_trace(tmp)
    ev = json.load(open(tmp))["traceEvents"]
    os.unlink(tmp)
    kern = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
    by = collections.defaultdict(lambda: [0, 0.0])
    for e in kern:
        n = e["name"].replace("void ", "")[:100]
        by[n][0] += 1
        by[n][1] += e["dur"]
    tot = sum(v[1] for v in by.values())
    prn = sum(v[1] for k, v in by.items() if k.startswith("prn::"))
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in kern)
    busy, cs, ce = 0.0, iv[0][0], iv[0][1]
    for s_, e_ in iv[1:]:
        if s_ > ce:
            busy += ce - cs
            cs, ce = s_, e_
        else:
            ce = max(ce, e_)
    busy += ce - cs
    print(f"device: {len(kern) // reps} kernels + copies per step, summed durations {tot / reps / 1e3:.2f} ms per step (prn:: kernels "
          f"{prn / reps / 1e3:.2f} ms), device busy {busy / reps / 1e3:.2f} ms per step")
    for n, (c, d) in sorted(by.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"  {n:100s} {c // reps:5d} {d / reps / 1e3:8.3f} ms")
    pr = cProfile.Profile()
    pr.enable()
    loop(3, True)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(30)


if __name__ == "__main__":
    main()
