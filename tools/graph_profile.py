"""Per-kernel durations INSIDE the replayed CUDA graphs (warm caches, real neighbours), from CUPTI activity records via
torch.profiler — the complement of the ncu launch lists, whose per-launch times are cold-cache and serialised.

  python tools/graph_profile.py train|infer [out_prefix]

Writes <out_prefix>_kernels.txt (per kernel name: launches, total, share; then per (name, grid): launches, mean) and prints the
first rows.  Not a bench: the profiler adds a few percent of overhead per launch."""
import collections
import json
import os
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from planerecnet_b200.config import cfg, set_cfg  # noqa: E402
from planerecnet_b200.planerecnet import PlaneRecNet  # noqa: E402
from planerecnet_b200.utils.synth import make_cotangents, make_input, perturb_  # noqa: E402


def short(name):
    name = name.replace("void ", "").replace("prn::", "")
    return name[:70]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "train"
    prefix = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "r02_graph_profile_" + mode)
    reps = 3
    set_cfg("PlaneRecNet_101_config")
    torch.manual_seed(0)
    net = perturb_(PlaneRecNet(cfg))
    x = make_input(8, 480, 640, 0).cuda()
    if mode == "train":
        from planerecnet_b200.train_engine import GraphedStep
        net = net.train().cuda()
        step = GraphedStep(net.train_engine, net, x)
        outs = step.forward(x)
        cots = make_cotangents(outs, seed=1, device="cuda")

        def run():
            step.forward(x)
            step.backward(*cots)
    else:
        net = net.eval().cuda()
        net.set_precision("bf16")
        with torch.no_grad():
            net.engine.forward_dense_graph(net, x, False)

        def run():
            with torch.no_grad():
                net.engine.forward_dense_graph(net, x, False)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(reps):
            run()
        torch.cuda.synchronize()
    tmp = tempfile.mktemp(suffix=".json")
    prof.export_chrome_trace(tmp)
    ev = json.load(open(tmp))["traceEvents"]
    os.unlink(tmp)
    kern = [e for e in ev if e.get("cat") == "kernel"]
    by_name = collections.defaultdict(lambda: [0, 0.0])
    by_grid = collections.defaultdict(lambda: [0, 0.0])
    for e in kern:
        n = short(e["name"])
        g = tuple(e.get("args", {}).get("grid", []))
        by_name[n][0] += 1
        by_name[n][1] += e["dur"]
        by_grid[(n, g)][0] += 1
        by_grid[(n, g)][1] += e["dur"]
    total = sum(v[1] for v in by_name.values())
    t0 = min(e["ts"] for e in kern)
    t1 = max(e["ts"] + e["dur"] for e in kern)
    lines = [f"# {mode}: {len(kern) // reps} kernel launches per replayed step, sum of kernel durations {total / reps / 1e3:.3f} ms per step, "
             f"span {(t1 - t0) / reps / 1e3:.3f} ms per step (includes the gaps between the {reps} replays)",
             f"{'kernel':72s} {'launches':>8s} {'ms/step':>9s} {'share':>7s}"]
    for n, (c, d) in sorted(by_name.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{n:72s} {c // reps:8d} {d / reps / 1e3:9.3f} {100 * d / total:6.1f}%")
    # timeline: how much of the span has no kernel running at all, and how busy each stream is
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in kern)
    busy, cur_s, cur_e, gaps = 0.0, iv[0][0], iv[0][1], []
    for s_, e_ in iv[1:]:
        if s_ > cur_e:
            busy += cur_e - cur_s
            gaps.append(s_ - cur_e)
            cur_s, cur_e = s_, e_
        else:
            cur_e = max(cur_e, e_)
    busy += cur_e - cur_s
    small = [g for g in gaps if g < 50]
    lines.append("")
    lines.append(f"# timeline per step: some kernel running {busy / reps / 1e3:.3f} ms; no kernel running (gaps < 50 us, i.e. inside a replay) "
                 f"{sum(small) / reps / 1e3:.3f} ms in {len(small) // reps} gaps (mean {sum(small) / max(len(small), 1):.2f} us)")
    per_stream = collections.defaultdict(float)
    for e in kern:
        per_stream[e.get("args", {}).get("stream", -1)] += e["dur"]
    lines.append("# kernel time per stream (ms/step): " + ", ".join(f"{k}: {v / reps / 1e3:.3f}" for k, v in sorted(per_stream.items(), key=lambda kv: -kv[1])))
    lines.append("")
    lines.append(f"{'kernel, grid':92s} {'launches':>8s} {'mean us':>9s} {'ms/step':>9s}")
    for (n, g), (c, d) in sorted(by_grid.items(), key=lambda kv: -kv[1][1])[:90]:
        lines.append(f"{n + ' ' + str(g):92s} {c // reps:8d} {d / c:9.1f} {d / reps / 1e3:9.3f}")
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    with open(prefix + "_kernels.txt", "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))
    print("\n".join(l for l in lines if l.startswith("# ")))


if __name__ == "__main__":
    main()
