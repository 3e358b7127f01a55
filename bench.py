#!/usr/bin/env python
"""Benchmark of the PlaneRecNet dense hot path on B200 (contract: see the task statement / DESIGN.md §5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--preset PlaneRecNet_101_config] [--batch 8] [--precision f16|bf16]

One "step" = one pass of the dense forward (backbone -> FPN -> instance/mask heads -> plane-prior
attention + depth decoder) over one synthetic batch of `--batch` 480x640 images per GPU.
`value`  : images/s, inputs resident in HBM, CUDA-graph replay, device-timed (CUDA events), max over ranks.
`e2e`    : images/s through the public module API `net(x)` (eval mode, incl. inference bookkeeping) with
           pinned HOST input, H2D inside the timed region and D2H of the detections + depth maps.
`--impl reference`: the reference algorithm's CPU path (oracle/prn_oracle.py, the CPU restatement pinned
           to the unmodified reference) on the host cores, bounded to one image per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec 480x640 ResNet101-DCN dense forward"
H_IMG, W_IMG = 480, 640
# algorithmic GFLOP per image of the conv-like contractions as the reference executes them
ALGO_GF = {"PlaneRecNet_101_config": 294.31, "PlaneRecNet_50_config": 249.30}
# training step fwd+bwd: as the reference executes it / as we execute it (the plane-prior attention is
# evaluated only at the pixels its x0.25 resize reads, forward and weight gradient alike)
ALGO_GF_TRAIN = {"PlaneRecNet_101_config": 808.2, "PlaneRecNet_50_config": 673.2}
EXEC_GF_TRAIN = {"PlaneRecNet_101_config": 725.7, "PlaneRecNet_50_config": 590.7}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="PlaneRecNet_101_config")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step (fwd+bwd) measurement")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return d.get("bf16_tflops_sustained", 1374.9), d.get("hbm_gbs", 6553.6), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 7 for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        mx = max([int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()] or [0])
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


def cpu_reference(preset, steps, warmup):
    """The reference algorithm on the host cores: oracle forward_eval (dense forward + inference bookkeeping)
    of ONE 480x640 image per step (bounded sample of the bs=8 workload)."""
    import torch
    from oracle import prn_oracle as O
    from planerecnet_b200.config import cfg, set_cfg
    from planerecnet_b200.planerecnet import PlaneRecNet
    from planerecnet_b200.utils.synth import make_input, perturb_
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    set_cfg(preset)
    torch.manual_seed(0)
    net = perturb_(PlaneRecNet(cfg)).eval()
    orc = O.Oracle(net.state_dict(), preset)
    x = make_input(1, H_IMG, W_IMG, 0)
    with torch.no_grad():
        for _ in range(warmup):
            orc.forward_eval(x)
        t0 = time.perf_counter()
        for _ in range(steps):
            orc.forward_eval(x)
        dt = time.perf_counter() - t0
    return {"value": steps / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps x 1 image 480x640 ({preset}), oracle forward_eval, fp32 torch CPU ops, {cores} threads",
            "ms_per_step": dt / steps * 1e3}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return 0
        steps, warm = min(a.steps, 6), min(a.warmup, 1)
        r = cpu_reference(a.preset, steps, warm)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": a.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{a.preset} eval forward 480x640 on host CPU, 1 image per step (bounded sample of bs={a.batch})"},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from planerecnet_b200.config import cfg, set_cfg
    from planerecnet_b200.planerecnet import PlaneRecNet
    from planerecnet_b200.utils.synth import make_input, perturb_

    set_cfg(a.preset)
    torch.manual_seed(0)
    net = perturb_(PlaneRecNet(cfg)).eval().cuda()
    net.set_precision(a.precision)
    eng = net.engine
    B = a.batch
    x_host = make_input(B, H_IMG, W_IMG, seed=rank).pin_memory()
    x_dev = x_host.cuda()

    from planerecnet_b200.utils import dist as D

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        return D.max_over_ranks(ms, device="cuda")     # planerecnet_b200/utils/dist.py (gloo-tested in tests/test_dist_cpu.py)

    # ---------------------------------------------------------------- device-resident throughput
    with torch.no_grad():
        for _ in range(max(a.warmup, 3)):
            eng.forward_dense_graph(net, x_dev, True)
        sync_all()
        sampler = ClockSampler(local)
        sampler.start()
        time.sleep(0.3)
        l0 = eng.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        e0.record()
        for _ in range(a.steps):
            eng.forward_dense_graph(net, x_dev, True)
        e1.record()
        sync_all()
        ms = max_over_ranks(e0.elapsed_time(e1))
        launches = eng.launches - l0
        clocks = sampler.stop()
    value = world * B * a.steps / (ms / 1e3)

    # ---------------------------------------------------------------- end to end through the public API
    # Serving loop through the public API `net.infer_pipelined(batches)`: every batch goes pinned host memory -> H2D
    # (copy stream) -> dense forward (graph replay) -> inference bookkeeping -> D2H of the detections + depth maps, all
    # inside the timed region; batch k+1's copy and forward overlap batch k's bookkeeping (two graph slots).
    pinned = {}

    def d2h(res):
        out_bytes = 0
        # device -> host read of the step's result: detections (scores, classes, boxes) and depth maps of all images,
        # concatenated per field (4 copies per step instead of 4 per image) into pinned host buffers
        staged = []
        for k in ("pred_scores", "pred_classes", "pred_boxes", "pred_depth"):
            parts = [r[k] for r in res if r[k] is not None]
            if parts:
                t = torch.cat(parts)
                buf = pinned.get(k)
                if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
                    buf = pinned[k] = torch.empty(max(t.numel(), 4096), dtype=t.dtype).pin_memory()
                dst = buf[:t.numel()]
                dst.copy_(t.reshape(-1), non_blocking=True)
                staged.append(dst)
                out_bytes += t.numel() * t.element_size()
        torch.cuda.current_stream().synchronize()      # the results are on the host when the step ends
        return out_bytes

    e_steps = max(3, min(a.steps, 10))
    e_warm = 5
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    d2h_bytes = 0
    # steady state at both ends of the timed window: t0 / t1 are recorded right after a result has been delivered while the
    # next batch's forward and the one after's copy are already in flight; two extra batches are fed to keep it so at the end
    for i, res in enumerate(net.infer_pipelined(x_host for _ in range(e_warm + e_steps + 3))):
        d2h_bytes = d2h(res)
        if i == e_warm - 1:
            t0.record()
        if i == e_warm + e_steps - 1:
            t1.record()
    sync_all()
    d2h = d2h_bytes
    e2e_ms = max_over_ranks(t0.elapsed_time(t1))
    e2e_value = world * B * e_steps / (e2e_ms / 1e3)

    # ---------------------------------------------------------------- roofline of the dominant kernel
    # conv_umma_kernel (all conv-like contractions): per-launch CUDA events on the launching stream over one
    # eager (un-graphed) step after warm-up; achieved = algorithmic FLOPs / summed kernel time.
    roof = None
    if rank == 0:
        with torch.no_grad():
            ms_flag, eng.multi_stream = eng.multi_stream, False      # one stream: event pairs bracket exactly one kernel
            eng.forward_dense(net, x_dev, False)
            torch.cuda.synchronize()
            eng.profile = []
            torch.cuda._sleep(int(6e7))        # ~30 ms GPU-side delay: the host enqueues the whole step behind it, so the
            eng.forward_dense(net, x_dev, False)   # event intervals measure kernel time, not Python launch latency
            torch.cuda.synchronize()
            prof, eng.profile = eng.profile, None
            eng.multi_stream = ms_flag
        by = {}
        for name, fl, s, e in prof:
            d = by.setdefault(name, [0.0, 0.0, 0])
            d[0] += fl
            d[1] += s.elapsed_time(e)
            d[2] += 1
        tot_f = sum(v[0] for v in by.values())
        tot_ms = sum(v[1] for v in by.values())
        peak, hbm, how = peaks()
        ach = tot_f / (tot_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "conv_umma_kernel", "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s",
                "frac": round(ach / peak, 4), "traffic": None, "peak_source": f"{how} bf16_tflops_sustained",
                "launches_per_step": sum(v[2] for v in by.values()), "kernel_ms_per_step": round(tot_ms, 3),
                "algorithmic_gflop_per_step": round(tot_f / 1e9, 1),
                "graph_step_frac": round(tot_f / (ms / a.steps / 1e3) / 1e12 / peak, 4),
                "by_kind": {k: {"gflop": round(v[0] / 1e9, 1), "ms": round(v[1], 3), "launches": v[2],
                                "tflops": round(v[0] / (v[1] / 1e3) / 1e12, 1) if v[1] > 0 else None}
                            for k, v in sorted(by.items())}}

    # ---------------------------------------------------------------- training step: fwd + bwd (+ gradient all-reduce)
    # net.train() (batch-statistics BatchNorm), model forward + backward driven by fixed seeded
    # cotangents of the 10 output tensors, bf16 activations/gradients, fp32 weight gradients; the step replays two
    # captured CUDA graphs (weight packing included).  With N > 1 ranks every step ends with ONE NCCL all-reduce (mean)
    # over the flat gradient buffer (§8e).
    train = None
    if not a.no_train:
        try:
            from planerecnet_b200.train_engine import GraphedStep
            torch.manual_seed(0)
            tnet = perturb_(PlaneRecNet(cfg)).train().cuda()
            teng = tnet.train_engine
            from planerecnet_b200.optim import FusedAdam
            lr = 1e-4          # train.py:251-256: Adam, five parameter groups
            opt = FusedAdam([{"params": list(tnet.backbone.parameters()), "lr": 5 * lr}, {"params": list(tnet.fpn.parameters()), "lr": lr},
                             {"params": list(tnet.inst_head.parameters()), "lr": lr}, {"params": list(tnet.mask_head.parameters()), "lr": lr},
                             {"params": list(tnet.depth_decoder.parameters()), "lr": 2 * lr}], lr=lr)
            step = GraphedStep(teng, tnet, x_dev, optimizer=opt, world=world)
            gen = torch.Generator(device="cuda").manual_seed(1 + rank)

            def mk(t):
                return torch.randn(t.shape, device="cuda", generator=gen) / t[0].numel() ** 0.5

            m_, cs_, ks_, d_ = step.outs
            cots = (mk(m_), [mk(c) for c in cs_], [mk(k) for k in ks_], mk(d_))
            params = [p for p in tnet.parameters() if p.requires_grad]

            def train_step():
                step.forward(x_dev)
                g = step.backward(*cots)
                if world > 1:
                    D.allreduce_mean_grads(g, params)      # one NCCL all-reduce over the flat gradient buffer (gloo-tested on CPU)
                return g

            for _ in range(3):
                train_step()
            sync_all()
            t_steps = max(3, min(a.steps, 10))
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            ev[0].record()
            for _ in range(t_steps):
                train_step()
            ev[1].record()
            sync_all()
            t_ms = max_over_ranks(ev[0].elapsed_time(ev[1])) / t_steps
            ev[0].record()
            step.forward(x_dev)
            ev[1].record()
            step.backward(*cots)
            ev[2].record()
            torch.cuda.synchronize()
            fwd_ms, bwd_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
            # full iteration: forward + backward + (gradient all-reduce) + Adam update, all replayed from graphs
            it_ms, it_err = None, None
            try:
                for _ in range(2):
                    step.forward(x_dev)
                    step.backward(*cots)
                    step.optimizer_step()
                sync_all()
                ev[0].record()
                for _ in range(t_steps):
                    step.forward(x_dev)
                    step.backward(*cots)
                    step.optimizer_step()
                ev[1].record()
                sync_all()
                it_ms = max_over_ranks(ev[0].elapsed_time(ev[1])) / t_steps
            except Exception as exc:
                it_err = f"{type(exc).__name__}: {exc}"[:300]
            peak, _, how = peaks()
            ex = EXEC_GF_TRAIN.get(a.preset)
            train = {"value": round(world * B / (t_ms / 1e3), 2), "unit": "images/s", "ms_per_step": round(t_ms, 3),
                     "fwd_ms": round(fwd_ms, 3), "bwd_ms": round(bwd_ms, 3), "steps": t_steps,
                     "dtype": "bf16", "gpu_launches_per_step": step.fwd_launches + step.bwd_launches,
                     "grad_allreduce": "nccl, 1 flat fp32 buffer / step" if world > 1 else None,
                     "algorithmic_gflop_per_image": ALGO_GF_TRAIN.get(a.preset), "executed_gflop_per_image": ex,
                     "tensor_frac_of_peak": round(ex * B / t_ms / peak, 4) if ex else None,
                     "full_iteration": ({"error": it_err} if it_ms is None else
                                        {"value": round(world * B / (it_ms / 1e3), 2), "unit": "images/s", "ms_per_step": round(it_ms, 3),
                                         "what": "fwd + bwd + " + ("NCCL all-reduce of the flat gradient buffer + " if world > 1 else "") +
                                                 "fused Adam over 5 parameter groups (prn_adam_multi), 3 graph replays"}),
                     "workload": f"{a.preset} net.train() fwd+bwd bs={B}/GPU 480x640, fixed seeded cotangents (kernel-only step, "
                                 f"no loss/optimizer), 2 CUDA graphs incl. weight packing"}
            del step, tnet
            torch.cuda.empty_cache()

        except Exception as exc:      # never lose the headline line to the secondary measurement
            train = {"error": f"{type(exc).__name__}: {exc}"[:300]}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference(a.preset, 2, 1)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": a.steps,
                "warmup": max(a.warmup, 3), "ms_per_step": round(ms / a.steps, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": a.precision, "data": "synthetic",
                "config": {"workload": f"{a.preset} inference bs={B}/GPU 480x640: dense forward backbone->FPN->heads->depth, "
                                       f"random-init weights, CUDA-graph replay",
                           "global_batch": world * B, "parallelism": f"dp{world} (replicas, no data-path collective)",
                           "l2": "no explicit flush: one step streams >3 GB of activations through a 126 MB L2",
                           "algorithmic_gflop_per_image": ALGO_GF.get(a.preset)},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": round(e2e_value, 2), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
                        "d2h_bytes_per_step": d2h, "steps": e_steps,
                        "path": "net.infer_pipelined(batches): pinned host input -> H2D (copy stream) -> graph forward -> inference bookkeeping -> D2H; batch k+1's copy + forward overlap batch k's bookkeeping"},
                "roofline": roof, "cpu_baseline": cpu, "train_step": train}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
