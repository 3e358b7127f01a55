#!/usr/bin/env python
"""Benchmark of the PlaneRecNet dense hot path on B200 (contract: the task statement / DESIGN.md §5).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                  [--preset PlaneRecNet_101_config] [--batch 8] [--precision f16|bf16]

Headline (BASELINE.json metric: "images/sec 480x640 ResNet101-DCN fwd+bwd at 1/2/4/8 B200"):
one "step" = one training step of PlaneRecNet_101_config over one synthetic batch of `--batch` 480x640 images per GPU:
net.train() forward (batch-statistics BatchNorm) + backward of every parameter, and with N > 1 ranks the gradient
all-reduce (mean) over NCCL.
  value   : images/s of that step with inputs resident in HBM, backward driven by fixed seeded cotangents of the 10
            output tensors (config 4 (i), the kernel-only step the 808.2 GF/image figure describes), CUDA-graph replay,
            device-timed (CUDA events), max over ranks.
  e2e     : images/s of the loop a user runs through the public API (config 4 (ii)), with one batch of look-ahead like a
            prefetching DataLoader: pinned HOST images + ground truth -> H2D (copy stream) -> `crit.prepare(gts)` -> `net(x)` ->
            `PlaneRecNetLoss` -> `loss.backward()` (+ all-reduce) -> D2H of the five loss terms + stream sync, every step;
            `no_lookahead` and `device_sampling` variants beside it.
  inference: the eval-mode dense forward (configs 2 / 3; north_star's >= 70 % tensor-pipe target is quoted on it) with its
            own value / e2e (`net.infer_pipelined`) and the per-kind conv roofline.
  gpu_reference: the UNMODIFIED reference (baseline/_ref) on the same B200 through cuDNN + torchvision deform_conv2d:
            the kernel set to beat.
`--impl reference`: the UNMODIFIED reference's CPU path (baseline/ref_runner.py) on the host cores, same training step
            (forward + PlaneRecNetLoss + backward), bounded to one image per step.
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
# more hardware work queues than the default 8: the step uses ~10 streams (side-stream weight gradients, decoder / head branches,
# copy / forward / bookkeeping streams of the serving loop); streams that alias onto one queue serialise falsely
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

# stdout carries exactly ONE line, the JSON: everything else that writes to file descriptor 1 (NCCL's version banner under torchrun,
# library chatter) is sent to stderr, and the line itself goes to a duplicate of the original stdout
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


METRIC = "images/sec 480x640 ResNet101-DCN fwd+bwd"
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
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16"], help="inference operand type")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-inference", action="store_true", help="skip the eval-mode forward measurements")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-bench parity check against the oracle")
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


def run_ref(args, timeout=900):
    """baseline/ref_runner.py in its own process (the unmodified reference's module names never mix with ours)."""
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_runner.py")] + args
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"unavailable": f"ref_runner timed out after {timeout} s"}
    for line in reversed(out.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    return {"unavailable": ("ref_runner failed: " + out.stderr.strip()[-300:]) if out.stderr else "ref_runner printed nothing"}


def cpu_reference_arm(preset, steps, warmup):
    """The reference's own CPU implementation of the benchmarked step on the host cores: UNMODIFIED reference (baseline/_ref)
    net.train() forward + PlaneRecNetLoss + backward of ONE 480x640 image per step (bounded sample of the bs = 8 workload)."""
    r = run_ref(["--device", "cpu", "--mode", "train_loss", "--preset", preset, "--batch", "1", "--steps", str(steps),
                 "--warmup", str(warmup)])
    if "unavailable" in r:
        return cpu_oracle_port(preset, steps, warmup, r["unavailable"])
    return {"value": r["images_per_s"], "unit": "images/s", "cores": r["cores"], "kind": "reference",
            "sample": f"{steps} steps x 1 image 480x640 ({preset}): unmodified reference (baseline/_ref, manifest ok: {r['manifest_ok']}) "
                      f"net.train() forward + PlaneRecNetLoss + backward, fp32 torch CPU ops, {r['cores']} threads",
            "ms_per_step": r["ms_per_step"]}


def cpu_oracle_port(preset, steps, warmup, why):
    """Fallback when baseline/_ref is not staged: the oracle (a port pinned to the reference), forward + backward."""
    import torch
    from oracle import prn_oracle as O
    from planerecnet_b200.config import cfg, set_cfg
    from planerecnet_b200.planerecnet import PlaneRecNet
    from planerecnet_b200.utils.synth import make_input, perturb_
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    set_cfg(preset)
    torch.manual_seed(0)
    net = perturb_(PlaneRecNet(cfg)).train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = make_input(1, H_IMG, W_IMG, 0)

    def step():
        o = O.Oracle(sd, preset, bn_train=True)
        for k, v in o.sd.items():
            if v.is_floating_point() and "running" not in k:
                v.requires_grad_(True)
        m, cs, ks, d = o.forward_dense(x)
        sum((a * a).sum() for a in [m] + list(cs) + list(ks) + [d]).backward()

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": steps / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps x 1 image 480x640 ({preset}), oracle forward + backward (port; reference not staged: {why}), "
                      f"fp32 torch CPU ops, {cores} threads", "ms_per_step": dt / steps * 1e3}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        if rank != 0:
            return 0
        steps, warm = min(a.steps, 2), min(a.warmup, 1)
        r = cpu_reference_arm(a.preset, steps, warm)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "images/s", "n_gpus": a.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{a.preset} train step (forward + PlaneRecNetLoss + backward) 480x640 on the host CPU, "
                                       f"1 image per step (bounded sample of bs={a.batch})"},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ---------------------------------------------------------------- the reference on the same B200 (own process, first:
    # the GPU is otherwise idle).  fp32 with cuDNN's TF32 default = what train.py / eval.py run; fp16 autocast + channels_last
    # = the fastest stock configuration (torchvision's deform_conv2d stays fp32 under autocast).
    gpu_ref = None
    if rank == 0 and world == 1 and not a.no_gpu_reference:
        gpu_ref = {}
        for key, extra in (("fwd_dense", []), ("train_cot", []), ("train_loss", []), ("fwd_dense_amp", ["--amp", "--channels-last"]),
                           ("train_cot_amp", ["--amp", "--channels-last"])):
            mode = key.replace("_amp", "")
            r = run_ref(["--device", "cuda", "--mode", mode, "--preset", a.preset, "--batch", str(a.batch), "--steps", "8",
                         "--warmup", "3"] + extra, timeout=300)
            gpu_ref[key] = ({"unavailable": r["unavailable"]} if "unavailable" in r else
                            {"value": round(r["images_per_s"], 2), "unit": "images/s", "ms_per_step": round(r["ms_per_step"], 3),
                             "dtype": r["dtype"], "channels_last": r["channels_last"]})
        gpu_ref["what"] = ("unmodified reference (baseline/_ref) on this GPU: cuDNN convolutions + torchvision CUDA deform_conv2d, eager "
                           "PyTorch, same preset / batch / synthetic input; fwd_dense = eval-BN dense forward, train_cot = net.train() "
                           "fwd+bwd with seeded cotangents (config 4 (i)), train_loss = fwd + PlaneRecNetLoss + bwd (config 4 (ii))")

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from planerecnet_b200.config import cfg, set_cfg
    from planerecnet_b200.planerecnet import PlaneRecNet
    from planerecnet_b200.utils import dist as D
    from planerecnet_b200.utils.synth import make_cotangents, make_gt, make_input, perturb_

    set_cfg(a.preset)
    B = a.batch
    x_host = make_input(B, H_IMG, W_IMG, seed=rank).pin_memory()
    x_dev = x_host.cuda()
    peak, hbm, how = peaks()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(ms):
        return D.max_over_ranks(ms, device="cuda")     # planerecnet_b200/utils/dist.py (gloo-tested in tests/test_dist_cpu.py)

    # ================================================================ headline: training step fwd + bwd (+ all-reduce)
    from planerecnet_b200.optim import FusedAdam
    from planerecnet_b200.train_engine import GraphedStep
    torch.manual_seed(0)
    tnet = perturb_(PlaneRecNet(cfg)).train().cuda()
    teng = tnet.train_engine
    lr = 1e-4          # train.py:251-256: Adam, five parameter groups
    opt = FusedAdam([{"params": list(tnet.backbone.parameters()), "lr": 5 * lr}, {"params": list(tnet.fpn.parameters()), "lr": lr},
                     {"params": list(tnet.inst_head.parameters()), "lr": lr}, {"params": list(tnet.mask_head.parameters()), "lr": lr},
                     {"params": list(tnet.depth_decoder.parameters()), "lr": 2 * lr}], lr=lr)
    step = GraphedStep(teng, tnet, x_dev, optimizer=opt, world=world)
    cots = make_cotangents(step.outs, seed=1 + rank, device="cuda")

    def train_step():
        step.forward(x_dev)
        g = step.backward(*cots)
        if world > 1:
            step.allreduce_grads()      # gradient all-reduce (mean) over NCCL, SURVEY §8e
        return g

    W = max(a.warmup, 3)
    for _ in range(W):
        train_step()
    sync_all()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = teng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record()
    for _ in range(a.steps):
        train_step()
    e1.record()
    sync_all()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = teng.launches - l0
    clocks = sampler.stop()
    value = world * B * a.steps / (ms / 1e3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record()
    step.forward(x_dev)
    ev[1].record()
    step.backward(*cots)
    ev[2].record()
    torch.cuda.synchronize()
    fwd_ms, bwd_ms = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    # full iteration: forward + backward + (gradient all-reduce) + Adam update, all replayed from graphs
    it_ms, it_err = None, None
    t_steps = max(3, min(a.steps, 10))
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
    ex = EXEC_GF_TRAIN.get(a.preset)
    ms_step = ms / a.steps
    train_detail = {"fwd_ms": round(fwd_ms, 3), "bwd_ms": round(bwd_ms, 3), "dtype": "bf16",
                    "gpu_launches_per_step": step.fwd_launches + step.bwd_launches,
                    "grad_allreduce": step.allreduce_desc() if world > 1 else None,
                    "algorithmic_gflop_per_image": ALGO_GF_TRAIN.get(a.preset), "executed_gflop_per_image": ex,
                    "tensor_frac_of_peak": round(ex * B / ms_step / peak, 4) if ex else None,
                    "full_iteration": ({"error": it_err} if it_ms is None else
                                       {"value": round(world * B / (it_ms / 1e3), 2), "unit": "images/s", "ms_per_step": round(it_ms, 3),
                                        "what": "fwd + bwd + " + ("NCCL all-reduce of the gradients + " if world > 1 else "") +
                                                "fused Adam over 5 parameter groups (prn_adam_multi), graph replays"})}
    del step
    torch.cuda.empty_cache()

    # ================================================================ e2e: the step a user runs (config 4 (ii))
    # pinned host images + ground truth -> H2D -> net(x) (graphed forward) -> PlaneRecNetLoss -> backward (graphed) ->
    # (all-reduce) -> D2H of the loss terms.  Everything inside the timed region, every step.
    e2e = None
    try:
        import numpy as np
        from planerecnet_b200.losses import PlaneRecNetLoss
        crit = PlaneRecNetLoss(cfg)
        gts_h, gtd_h = make_gt(B, H_IMG, W_IMG, seed=rank)
        gts_h = [{k: v.pin_memory() for k, v in g.items()} for g in gts_h]
        gtd_h = gtd_h.pin_memory()
        h2d = x_host.numel() * 4 + gtd_h.numel() * 4 + sum(v.numel() * v.element_size() for g in gts_h for v in g.values())
        tnet.use_train_graph = True
        params = [p for p in tnet.parameters() if p.requires_grad]
        loss_host = torch.empty(5).pin_memory()

        copy_stream = torch.cuda.Stream()
        # the loss preparation of the NEXT batch runs on worker threads while this thread launches the backward: with CPython's
        # default 5 ms switch interval every GIL hand-over can stall the launching thread for up to 5 ms (measured: 37 -> 54 ms
        # per step on an unlucky run); 0.2 ms keeps the hand-overs out of the way.  Restored after the e2e measurement.
        switch_interval = sys.getswitchinterval()
        sys.setswitchinterval(2e-4)

        def upload():
            """One batch: pinned host -> device on the copy stream, and the ground-truth-only part of its loss started."""
            with torch.cuda.stream(copy_stream):
                x = x_host.cuda(non_blocking=True)
                gts = [{k: v.cuda(non_blocking=True) for k, v in g.items()} for g in gts_h]
                gtd = gtd_h.cuda(non_blocking=True)
                crit.prepare(gts)              # targets + triplet sampling: worker thread + side stream, returns at once
            return x, gts, gtd

        def consume(batch):
            main = torch.cuda.current_stream()
            main.wait_stream(copy_stream)
            x, gts, gtd = batch
            for t in [x, gtd] + [v for g in gts for v in g.values()]:
                t.record_stream(main)
            return x, gts, gtd

        # where the next batch's upload + loss preparation start: after this step's backward has been launched (default: the
        # preparation threads then run while this thread only waits for the device; 33-35 ms per step, steady) or before it
        # (PRN_PREFETCH_EARLY=1: 32.7 ms at best, but the threads then compete with the backward's launches for the GIL and an
        # occasional run falls back to 45 ms)
        prefetch_early = os.environ.get("PRN_PREFETCH_EARLY") == "1"

        def step_body(x, gts, gtd, prefetch):
            for p in params:
                p.grad = None
            mask, cate, kern, depth = tnet(x)
            losses = crit(tnet, mask, cate, kern, depth, gts, gtd)
            losses = {k: v.mean() for k, v in losses.items()}          # train.py:347-348
            nxt = upload() if (prefetch and prefetch_early) else None   # next batch: H2D + its GT-only work overlap this backward
            sum(losses[k] for k in losses).backward()
            if world > 1:
                D.allreduce_mean_grads({id(p): p.grad for p in params if p.grad is not None}, params)
            if prefetch and not prefetch_early:
                nxt = upload()                                         # ... or only the device's share of it (host is done launching)
            vals = torch.stack([losses[k].detach().float() for k in ("ins", "cat", "dpt", "pln", "lav")])
            loss_host.copy_(vals, non_blocking=True)
            # the losses are on the host when the step ends.  A blocking event instead of stream.synchronize(): the latter spins on
            # a core for the ~15 ms the GPU still needs, next to the sampler / preparation threads of the next batch
            if os.environ.get("PRN_BLOCKING_SYNC") == "1":
                done = torch.cuda.Event(blocking=True)
                done.record()
                done.synchronize()
            else:
                torch.cuda.current_stream().synchronize()
            return loss_host.tolist(), nxt

        def user_loop(n, lookahead):
            """n steps of the loop a user writes.  lookahead: one batch of prefetch (the next batch's upload and the
            ground-truth-only part of its loss are started while this step's backward runs, like a prefetching DataLoader);
            otherwise every step uploads its own batch first.  n uploads and n loss read-backs either way."""
            np.random.seed(0)
            nxt = upload() if lookahead else None
            for i in range(n):
                batch = consume(nxt if lookahead else upload())
                lv_, nxt = step_body(*batch, prefetch=lookahead and i + 1 < n)
            return lv_

        def timed_user_steps(lookahead=True):
            user_loop(3, lookahead)
            sync_all()
            n = max(3, min(a.steps, 8))
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            lv_ = user_loop(n, lookahead)
            t1.record()
            sync_all()
            return max_over_ranks(t0.elapsed_time(t1)) / n, n, lv_

        u_ms, u_steps, lv = timed_user_steps()
        # the same step with the plane term's triplets drawn on the device (same distribution, no numpy RNG stream on the host):
        # the default above reproduces the reference's samples bit for bit (C restatement of numpy's legacy choice + shuffle on
        # numpy's own MT19937 state, ~25 ms per step on a worker thread, overlapped with the forward)
        s_ms, _, _ = timed_user_steps(lookahead=False)
        crit = PlaneRecNetLoss(cfg, vnl_sampling="device")
        f_ms, _, _ = timed_user_steps()
        sys.setswitchinterval(switch_interval)
        e2e = {"value": round(world * B / (u_ms / 1e3), 2), "unit": "images/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 20,
               "steps": u_steps, "ms_per_step": round(u_ms, 3),
               "no_lookahead": {"value": round(world * B / (s_ms / 1e3), 2), "unit": "images/s", "ms_per_step": round(s_ms, 3),
                                "what": "same loop without the one-batch look-ahead: every step uploads its own batch and starts its loss "
                                        "preparation right before its forward"},
               "device_sampling": {"value": round(world * B / (f_ms / 1e3), 2), "unit": "images/s", "ms_per_step": round(f_ms, 3),
                                   "what": "same step with PlaneRecNetLoss(vnl_sampling='device'): plane-term triplets from torch's device RNG "
                                           "(same distribution) instead of numpy's global RNG in the reference's call order"},
               "losses": {k: round(v, 5) for k, v in zip(("ins", "cat", "dpt", "pln", "lav"), lv)},
               "path": "training loop with one batch of look-ahead: pinned host images + ground truth -> H2D (copy stream) -> net(x) "
                       "[train mode, graphed fwd] -> planerecnet_b200.losses.PlaneRecNetLoss (crit.prepare(gts) at upload time: device-side target assignment + numpy-stream-exact triplet sampling on a worker thread; kernel-backed dice/lava/focal/depth terms, batched plane term) -> loss.backward() "
                       "[graphed bwd]" + (" -> NCCL all-reduce of the gradients" if world > 1 else "") + " -> D2H of the 5 loss terms"}
    except Exception as exc:
        e2e = {"value": None, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "error": f"{type(exc).__name__}: {exc}"[:300]}
    del tnet
    torch.cuda.empty_cache()

    # ================================================================ inference: eval-mode dense forward (configs 2 / 3)
    inference, roof, parity = None, None, None
    if not a.no_inference:
        torch.manual_seed(0)
        net = perturb_(PlaneRecNet(cfg)).eval().cuda()
        net.set_precision(a.precision)
        eng = net.engine
        with torch.no_grad():
            for _ in range(W):
                eng.forward_dense_graph(net, x_dev, True)
            sync_all()
            i0 = eng.launches
            e0.record()
            for _ in range(a.steps):
                eng.forward_dense_graph(net, x_dev, True)
            e1.record()
            sync_all()
            i_ms = max_over_ranks(e0.elapsed_time(e1))
            i_launches = eng.launches - i0
        i_value = world * B * a.steps / (i_ms / 1e3)

        # ---- parity of the timed batch: image 0 of this rank's batch against the CPU oracle (pinned to the reference)
        if rank == 0 and not a.no_parity:
            try:
                from oracle import prn_oracle as O
                with torch.no_grad():
                    st = eng.forward_dense_graph(net, x_dev, True)
                    got = [st["outputs"][0][:1].cpu(), st["outputs"][3][:1].cpu()] + [k[:1].cpu() for k in st["outputs"][2]]
                    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
                    om, oc, ok_, od = O.Oracle(sd, a.preset).forward_dense(x_host[:1])
                refs = [om, od] + list(ok_)

                def rel(p, q):
                    return float((p.double() - q.double()).norm() / (q.double().norm() + 1e-30))

                errs = [rel(g_, r_) for g_, r_ in zip(got, refs)]
                tol = 3e-3 if a.precision == "f16" else 2e-2
                parity = {"against": "oracle/prn_oracle.py (CPU fp32, pinned to the unmodified reference) on image 0 of the timed batch",
                          "rel_l2": {"mask": round(errs[0], 6), "depth": round(errs[1], 6), "kernel_pred_max": round(max(errs[2:]), 6)},
                          "tolerance": tol, "ok": bool(max(errs) <= tol)}
                assert parity["ok"], f"bench parity check failed: {parity}"
            except AssertionError:
                raise
            except Exception as exc:
                parity = {"error": f"{type(exc).__name__}: {exc}"[:300]}

        # ---- end to end through the public serving API
        from planerecnet_b200.postprocess import pack_mask_bits
        pinned = {}

        def d2h(res):
            """Device -> host read of the step's result: scores, classes, boxes, depth maps and the instance masks (bool
            [n, H, W] each, bit-packed on the device: 8 pixels per byte), concatenated per field into pinned host buffers."""
            out_bytes = 0
            staged = []
            fields = {k: [r[k] for r in res if r[k] is not None] for k in ("pred_scores", "pred_classes", "pred_boxes", "pred_depth")}
            packed = pack_mask_bits([r["pred_masks"] for r in res])            # prn_pack_mask_bits: one launch for the batch
            if packed is not None:
                fields["pred_masks_packed"] = [packed]
            for k, parts in fields.items():
                if parts:
                    t = torch.cat(parts) if len(parts) > 1 else parts[0]
                    buf = pinned.get(k)
                    if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
                        buf = pinned[k] = torch.empty(max(int(t.numel() * 1.5), 4096), dtype=t.dtype).pin_memory()
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
        for i, res in enumerate(net.infer_pipelined(x_host for _ in range(e_warm + e_steps + 3))):
            d2h_bytes = d2h(res)
            if i == e_warm - 1:
                t0.record()
            if i == e_warm + e_steps - 1:
                t1.record()
        sync_all()
        ie_ms = max_over_ranks(t0.elapsed_time(t1))
        ie_value = world * B * e_steps / (ie_ms / 1e3)

        # ---- roofline of the dominant kernel family: the conv-like contractions (conv_tma_kernel / conv_umma_kernel),
        # per-launch CUDA events on the launching stream over one eager (un-graphed, single-stream) step after warm-up
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
                dd = by.setdefault(name, [0.0, 0.0, 0])
                dd[0] += fl
                dd[1] += s.elapsed_time(e)
                dd[2] += 1
            tot_f = sum(v[0] for v in by.values())
            tot_ms = sum(v[1] for v in by.values())
            ach = tot_f / (tot_ms / 1e3) / 1e12
            traffic, traffic_detail = None, None
            tp = os.path.join(ROOT, "profiles", "r02_traffic.json")      # dram bytes per launch from the committed ncu captures
            if os.path.exists(tp):
                with open(tp) as fh:
                    traffic_detail = json.load(fh)
                top = traffic_detail["per_launch"].get("fpn0_3x3_256")    # the largest conv launch of the step (181 GFLOP)
                if top:
                    traffic = top["dram_bytes_read"] + top["dram_bytes_write"]
            roof = {"bound": "tensor", "kernel": "conv_tma_kernel + conv_umma_kernel (every conv-like contraction of the eval forward)",
                    "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s", "frac": round(ach / peak, 4),
                    "traffic": traffic, "traffic_what": "dram__bytes_read.sum + dram__bytes_write.sum of the largest conv launch (3x3 256->256 at "
                    "120x160, bs 8: 158.5 MB algorithmic in + out) from profiles/r02_ncu_full_summary.txt; other layers in traffic_detail",
                    "traffic_detail": traffic_detail, "peak_source": f"{how} bf16_tflops_sustained",
                    "launches_per_step": sum(v[2] for v in by.values()), "kernel_ms_per_step": round(tot_ms, 3),
                    "algorithmic_gflop_per_step": round(tot_f / 1e9, 1),
                    "graph_step_frac": round(tot_f / (i_ms / a.steps / 1e3) / 1e12 / peak, 4),
                    "train_step_frac": train_detail["tensor_frac_of_peak"],
                    "by_kind": {k: {"gflop": round(v[0] / 1e9, 1), "ms": round(v[1], 3), "launches": v[2],
                                    "tflops": round(v[0] / (v[1] / 1e3) / 1e12, 1) if v[1] > 0 else None}
                                for k, v in sorted(by.items())}}
        inference = {"value": round(i_value, 2), "unit": "images/s", "ms_per_step": round(i_ms / a.steps, 4), "dtype": a.precision,
                     "gpu_launches": i_launches, "algorithmic_gflop_per_image": ALGO_GF.get(a.preset),
                     "e2e": {"value": round(ie_value, 2), "unit": "images/s", "h2d_bytes_per_step": x_host.numel() * 4,
                             "d2h_bytes_per_step": d2h_bytes, "steps": e_steps, "masks": "bit-packed on the device, copied to the host",
                             "path": "net.infer_pipelined(batches): pinned host input -> H2D (copy stream) -> graph forward -> "
                                     "inference bookkeeping -> D2H of scores, classes, boxes, depth and bit-packed masks"},
                     "parity": parity,
                     "workload": f"{a.preset} eval forward bs={B}/GPU 480x640 (backbone->FPN->heads->plane-prior attention->depth), CUDA-graph replay"}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        r = cpu_reference_arm(a.preset, 1, 0)
        cpu = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        line = {"metric": METRIC, "value": round(value, 2), "unit": "images/s", "n_gpus": world, "steps": a.steps,
                "warmup": W, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": f"{a.preset} train step bs={B}/GPU 480x640: net.train() forward (batch-statistics BatchNorm) + backward "
                                       f"of all parameters driven by seeded cotangents of the 10 outputs (SURVEY §8d config 4 (i))"
                                       + (", gradient all-reduce (mean) over NCCL" if world > 1 else "") + ", random-init weights, CUDA-graph replay",
                           "global_batch": world * B, "parallelism": f"dp{world}" + (" (NCCL gradient all-reduce every step)" if world > 1 else ""),
                           "l2": "no explicit flush: one step streams >6 GB of activations through a 126 MB L2",
                           "algorithmic_gflop_per_image": ALGO_GF_TRAIN.get(a.preset)},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": e2e, "roofline": roof, "cpu_baseline": cpu, "train_step": train_detail, "inference": inference,
                "gpu_reference": gpu_ref}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
