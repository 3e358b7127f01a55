"""Training step through the public module API against autograd through the CPU oracle.

What is compared and why (tolerances are measured values with ~2x margin, stated per assert):
  * single bottlenecks with batch-statistics BatchNorm, identical inputs: every gradient, rel-L2;
  * whole model with frozen (running-statistics) BatchNorm: stable network -> tight gradient cosine; this pins the
    whole tape (every dgrad / wgrad / resampler / norm backward and every gradient join);
  * whole model with batch-statistics BatchNorm on a residual-dominant init: outputs and gradient cosine.  (On the
    plain random init the train-mode network is chaotic in the fp32 oracle itself — see train_cases._build.)
Gradients of a 16-bit training step are compared by cosine / rel-L2 (SURVEY §7 "hard parts" (c)), not element-wise."""
import pytest
import torch

import train_cases as TC


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["bf16", "f16"])
def test_bottleneck_blocks_train_bn(cuda_lib, prec):
    net = TC._build("PlaneRecNet_50_config").train()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    net = net.cuda()
    lim_all, lim_last, lim_out = (0.30, 0.05, 2.5e-2) if prec == "bf16" else (0.20, 8e-3, 4e-3)
    for (s, b) in ((0, 0), (0, 1), (1, 0), (1, 1), (2, 0), (3, 0), (3, 2)):     # plain / DCN, stride 1 / 2, +- projection
        r = TC.run_block_check(s, b, prec=prec, net=net, sd=sd)
        assert r["out"] <= lim_out, (s, b, r)
        assert r["conv3.weight"] <= lim_last and r["bn3.weight"] <= lim_last, (s, b, r)
        assert max(r.values()) <= lim_all, (s, b, r)


@pytest.mark.gpu
@pytest.mark.parametrize("preset,B", [("PlaneRecNet_50_config", 2), ("PlaneRecNet_101_config", 1)])
def test_model_step_frozen_bn(cuda_lib, preset, B):
    r = TC.run_model_check(preset, B=B, prec="bf16", bn_mode="frozen", cond=False)
    assert not r["missing"], r["missing"]
    assert max(r["outs"]) <= 2.5e-2, r["outs"]                 # bf16 forward, eval-BN: north_star's 1e-2-class bar
    assert r["all_cos"] >= 0.9995 and r["all_rel"] <= 0.03, (r["all_cos"], r["all_rel"])
    assert min(c for c, _ in r["fam"].values()) >= 0.985, r["fam"]


@pytest.mark.gpu
@pytest.mark.parametrize("prec", ["bf16", "f16"])
def test_model_step_train_bn(cuda_lib, prec):
    r = TC.run_model_check("PlaneRecNet_50_config", prec=prec, bn_mode="train", cond=True)
    assert not r["missing"], r["missing"]
    out_lim, cos_lim, bb_lim = (7e-2, 0.995, 0.70) if prec == "bf16" else (1e-2, 0.999, 0.92)
    assert max(r["outs"]) <= out_lim, r["outs"]
    assert r["all_cos"] >= cos_lim, r["all_cos"]
    for f in ("inst_head.cate_tower", "inst_head.kernel_tower", "mask_head.convs_all_levels", "fpn.fpn_convs"):
        assert r["fam"][f][0] >= 0.985, (f, r["fam"][f])
    assert r["fam"]["backbone.layers"][1] >= bb_lim, r["fam"]["backbone.layers"]
    # running statistics follow torch's update rule (momentum, unbiased variance), num_batches_tracked advanced
    st, sd0 = r["state"], r["sd0"]
    assert int(st["backbone.bn1.num_batches_tracked"]) == int(sd0["backbone.bn1.num_batches_tracked"]) + 1
    rm0, rm1 = sd0["backbone.bn1.running_mean"], st["backbone.bn1.running_mean"].cpu()
    assert float((rm1 - rm0).abs().max()) > 0


@pytest.mark.gpu
def test_running_stats_match_torch(cuda_lib):
    """backbone.bn1 after one training forward == nn.BatchNorm2d's update on the fp32 stem output."""
    from oracle import prn_oracle as O
    net = TC._build("PlaneRecNet_50_config").train()
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.randn(2, 3, 64, 96, generator=torch.Generator().manual_seed(3))
    netc = net.cuda()
    with torch.no_grad():
        netc(x.cuda())
    o = O.Oracle(sd0, "PlaneRecNet_50_config")
    y = o._conv(x, "backbone.conv1", 2, 3)
    rm, rv = sd0["backbone.bn1.running_mean"].clone(), sd0["backbone.bn1.running_var"].clone()
    torch.nn.functional.batch_norm(y, rm, rv, None, None, True, 0.1, 1e-5)
    got_m, got_v = netc.backbone.bn1.running_mean.cpu(), netc.backbone.bn1.running_var.cpu()
    assert float((got_m - rm).abs().max()) <= 2e-3 * float(rm.abs().max() + 1)
    assert float((got_v - rv).abs().max()) <= 5e-3 * float(rv.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("bn_mode", ["frozen", "train"])
def test_graphed_step_matches_eager_and_follows_optimizer(cuda_lib, bn_mode):
    """The two-graph replay of a step gives the eager step's gradients, and — because weight packing is captured too —
    keeps doing so after an in-place optimizer update.  With frozen statistics the network is stable and the two agree to
    fp32-atomics noise; with batch statistics run-to-run noise (atomic order -> 16-bit rounding flips, amplified by the
    train-mode network) is measured on the eager path itself and bounds the comparison."""
    from helpers import rel_l2
    net = TC._build("PlaneRecNet_50_config", cond=True).train()
    if bn_mode == "frozen":
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
    net = net.cuda()
    x = torch.randn(2, 3, 128, 160, generator=torch.Generator().manual_seed(5)).cuda()
    opt = torch.optim.SGD(net.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        m, cs, ks, d = net(x)
        outs = [m] + list(cs) + list(ks) + [d]
        sum((0.5 * a * a).sum() / a[0].numel() ** 0.5 for a in outs).backward()
        return torch.cat([p.grad.flatten().float() for p in net.parameters() if p.grad is not None]).clone()

    state = {k: v.clone() for k, v in net.state_dict().items()}
    net.use_train_graph = False
    g_eager0 = step()
    bn1 = net.backbone.bn1
    eager_bufs = (bn1.running_mean.clone(), bn1.running_var.clone(), int(bn1.num_batches_tracked))
    noise = rel_l2(step(), g_eager0)          # run-to-run noise floor of the eager path
    opt.step()
    g_eager1 = step()
    net.load_state_dict(state)
    net.use_train_graph = True
    g_graph0 = step()          # captures
    # the capture's eager warm-up must not count as a step: BatchNorm buffers after the first graphed step == after the
    # first eager step (running statistics once, num_batches_tracked + 1)
    nbt0 = int(state["backbone.bn1.num_batches_tracked"])
    if bn_mode == "train":
        assert eager_bufs[2] == nbt0 + 1 and int(bn1.num_batches_tracked) == nbt0 + 1, (eager_bufs[2], int(bn1.num_batches_tracked))
        for m_ in net.modules():
            if isinstance(m_, torch.nn.BatchNorm2d):
                assert int(m_.num_batches_tracked) == nbt0 + 1
        assert rel_l2(bn1.running_mean, eager_bufs[0]) < 2e-3 and rel_l2(bn1.running_var, eager_bufs[1]) < 2e-3
    else:
        assert int(bn1.num_batches_tracked) == nbt0 and torch.equal(bn1.running_mean, state["backbone.bn1.running_mean"])
    opt.step()
    g_graph1 = step()          # replays with the updated parameters
    tol = max(5 * noise, 1e-3)
    e0, e1, upd = rel_l2(g_graph0, g_eager0), rel_l2(g_graph1, g_eager1), rel_l2(g_eager1, g_eager0)
    print(f"bn={bn_mode} eager run-to-run {noise:.2e} graph-vs-eager {e0:.2e} {e1:.2e} change by the update {upd:.2e}")
    # after the update the two paths start from weights that already differ by the first step's noise, and the random-init
    # model amplifies it (the update changes the gradients by orders of magnitude): wider bound for the second comparison
    assert e0 <= tol and e1 <= 5 * tol, (e0, e1, noise)
    if bn_mode == "frozen":      # the update changed the gradients by more than the graph differs from eager: it was followed
        assert upd > 2 * e1, (upd, e1)


@pytest.mark.gpu
def test_fused_adam_matches_torch_adam(cuda_lib):
    """prn_adam_multi == torch.optim.Adam (train.py:251-256: parameter groups with their own learning rates), 4 steps,
    contiguous and 1-D strided gradients, tensors spanning several 65536-element chunks."""
    from planerecnet_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(0)
    shapes = [(300, 17, 3, 3), (70000,), (5,), (64, 64, 1, 1), (131072 + 7,)]
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    groups = lambda ps: [{"params": ps[:2], "lr": 5e-3}, {"params": ps[2:], "lr": 1e-3}]      # noqa: E731
    ref = torch.optim.Adam(groups(ref_p), lr=1e-3)
    ours = FusedAdam(groups(our_p), lr=1e-3)
    for it in range(4):
        grads = [torch.randn(s, generator=g).cuda() * (1 + it) for s in shapes]
        strided = torch.stack([grads[2], torch.zeros_like(grads[2])], 1)            # a [C,2] accumulator: column 0 is the gradient
        if it == 2:                       # set_lr (train.py:415-420) between steps
            for grp in ref.param_groups:
                grp["lr"] *= 0.5
            for grp in ours.param_groups:
                grp["lr"] *= 0.5
        for p, gr in zip(ref_p, grads):
            p.grad = gr.clone()
        ref.step()
        gd = {id(p): gr for p, gr in zip(our_p, grads)}
        gd[id(our_p[2])] = strided[:, 0]
        ours.step(gd)
    torch.cuda.synchronize()
    for a, b in zip(our_p, ref_p):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), float((a - b).abs().max())
    assert float(ours._state3[0]) == 4.0


@pytest.mark.gpu
def test_graphed_training_iteration_with_fused_adam(cuda_lib):
    """Forward, backward and the Adam update replayed from graphs track an eager torch.optim.Adam run on a twin model
    (frozen BatchNorm statistics: stable network).  Adam's first steps move every weight by ~lr * sign(gradient), so the
    comparison is on the distribution of parameter differences in units of lr."""
    from planerecnet_b200.optim import FusedAdam
    from planerecnet_b200.train_engine import GraphedStep

    def make():
        net = TC._build("PlaneRecNet_50_config", cond=False).train()
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.eval()
        return net.cuda()

    x = torch.randn(2, 3, 128, 160, generator=torch.Generator().manual_seed(5)).cuda()
    lr = 1e-4

    def groups(net):
        return [{"params": list(net.backbone.parameters()), "lr": 5 * lr}, {"params": list(net.fpn.parameters()), "lr": lr},
                {"params": list(net.inst_head.parameters()), "lr": lr}, {"params": list(net.mask_head.parameters()), "lr": lr},
                {"params": list(net.depth_decoder.parameters()), "lr": 2 * lr}]

    def cots_of(outs):
        m, cs, ks, d = outs
        flat = [m] + list(cs) + list(ks) + [d]
        c = [a.detach() / a[0].numel() ** 0.5 for a in flat]
        return c[0], c[1:1 + len(cs)], c[1 + len(cs):1 + 2 * len(cs)], c[-1]

    ref = make()
    opt_ref = torch.optim.Adam(groups(ref), lr=lr)
    for _ in range(3):
        opt_ref.zero_grad(set_to_none=True)
        m, cs, ks, d = ref(x)
        sum((0.5 * a * a).sum() / a[0].numel() ** 0.5 for a in [m] + list(cs) + list(ks) + [d]).backward()
        opt_ref.step()

    ours = make()
    p0 = [p.detach().clone() for p in ours.parameters()]
    opt = FusedAdam(groups(ours), lr=lr)
    step = GraphedStep(ours.train_engine, ours, x, optimizer=opt)
    for _ in range(3):
        outs = step.forward(x)
        step.backward(*cots_of(outs))
        step.optimizer_step()
    torch.cuda.synchronize()
    assert float(opt._state3[0]) == 3.0
    moved = torch.cat([(a.detach() - b).flatten() for a, b in zip(ours.parameters(), p0)])
    diff = torch.cat([(a.detach() - b.detach()).flatten() for a, b in zip(ours.parameters(), ref.parameters())])
    moved_ref = torch.cat([(a.detach() - b).flatten() for a, b in zip(ref.parameters(), p0)])
    cos = float((moved @ moved_ref) / (moved.norm() * moved_ref.norm()))
    print(f"adam: mean |update| {float(moved.abs().mean()):.3e}, mean |ours - torch| {float(diff.abs().mean()):.3e}, cosine {cos:.4f}")
    assert float(moved.abs().mean()) > 0.5 * lr                      # the graphs did update the weights
    # weights whose gradient is below the 16-bit noise floor get a +-lr step of arbitrary sign from Adam's normalisation
    assert cos > 0.9 and float(diff.abs().mean()) < 0.25 * float(moved.abs().mean()), (cos, float(diff.abs().mean()))


@pytest.mark.gpu
def test_f16_training_with_loss_scaling(cuda_lib):
    """set_train_precision('f16', grad_scale): tiny cotangents (5e-8-scale, below the f16 subnormal range after the first
    contraction) survive thanks to the scale, and the returned gradients are unscaled again."""
    from helpers import rel_l2
    net = TC._build("PlaneRecNet_50_config", cond=False).train()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    net = net.cuda()
    x = torch.randn(1, 3, 128, 160, generator=torch.Generator().manual_seed(7)).cuda()

    def grads(scale_loss):
        net.zero_grad(set_to_none=True)
        m, cs, ks, d = net(x)
        outs = [m] + list(cs) + list(ks) + [d]
        (scale_loss * sum((0.5 * a * a).sum() / a[0].numel() ** 0.5 for a in outs)).backward()
        return torch.cat([p.grad.flatten() for p in net.parameters() if p.grad is not None]).clone()

    net.set_train_precision("f16")
    g_ref = grads(1.0)
    net.set_train_precision("f16", grad_scale=2.0 ** 20)
    g_small = grads(1e-6)                   # a 1e-6-scaled loss: unscaled f16 gradients would underflow
    assert rel_l2(g_small, g_ref * 1e-6) < 5e-3, rel_l2(g_small, g_ref * 1e-6)


@pytest.mark.gpu
def test_bucketed_backward_graphs_equal_the_single_graph(cuda_lib):
    """GraphedStep(buckets=True) — the data-parallel form: backward captured as three graphs cut where a gradient bucket becomes
    final, each ending with the copy of its bucket into the flat all-reduce buffer — gives the single-graph gradients (frozen
    BatchNorm statistics: stable network), every parameter lands in exactly one bucket slice, and the flat views hold them."""
    from helpers import rel_l2
    from planerecnet_b200.train_engine import GraphedStep
    from planerecnet_b200.utils.synth import make_cotangents
    net = TC._build("PlaneRecNet_50_config", cond=True).train()
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.eval()
    net = net.cuda()
    x = torch.randn(2, 3, 128, 160, generator=torch.Generator().manual_seed(5)).cuda()
    eng = net.train_engine
    one = GraphedStep(eng, net, x)
    cots = make_cotangents(one.outs, seed=3, device="cuda")
    one.forward(x)
    g1 = {k: v.clone() for k, v in one.backward(*cots).items()}
    seg = GraphedStep(eng, net, x, buckets=True)
    seg.forward(x)
    seg.backward(*cots)
    views = seg.allreduce_grads()
    torch.cuda.synchronize()
    params = [p for p in net.parameters() if p.requires_grad and id(p) in g1]
    assert set(views.keys()) == {id(p) for p in params}
    assert sum(hi - lo for lo, hi in seg.bucket_slices) == seg.flat.numel() == sum(p.numel() for p in params)
    assert all(hi > lo for lo, hi in seg.bucket_slices)
    a = torch.cat([views[id(p)].flatten() for p in params])
    b = torch.cat([g1[id(p)].flatten().float() for p in params])
    assert bool(torch.isfinite(a).all())
    assert rel_l2(a, b) < 2e-2, rel_l2(a, b)


@pytest.mark.gpu
def test_repack_all_equals_the_per_conv_pack_kernels(cuda_lib):
    """TrainEngine.repack_all (one prn_pack_multi launch + one multi-tensor copy for every operand of the step) rewrites the
    remembered buffers with exactly what the per-conv pack kernels produce for the CURRENT weights."""
    from planerecnet_b200 import ops
    net = TC._build("PlaneRecNet_50_config").train().cuda()
    eng = net.train_engine
    x = torch.randn(1, 3, 128, 160, generator=torch.Generator().manual_seed(2)).cuda()
    outs = eng.forward_train(net, x)
    eng.seed_output_grads(torch.ones_like(outs[0]), [torch.ones_like(c) for c in outs[1]], [torch.ones_like(k) for k in outs[2]],
                          torch.ones_like(outs[3]))
    eng.backward()
    recs = eng._pack_recs
    assert len(recs) > 100 and {r["kind"] for r in recs.values()} == {0, 1}
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.25).add_(0.01)
    assert eng.repack_all(invalidate_others=False)
    torch.cuda.synchronize()
    for key, r in recs.items():
        w = r["w"].detach()
        if r["kind"] == 0:
            f = r["f"]
            splits = [(f[3], f[4])] + ([(f[6], f[0] - f[4])] if f[1] == 2 else [])
            ref = ops.pack_conv_weight_dev(w, splits, r["rows"], eng.dt)
            for param, buf in r["vec"]:
                assert torch.equal(buf[:param.numel()], param.detach().float()) and float(buf[param.numel():].abs().sum()) == 0.0
        else:
            lo, n, cout_pad = r["f"][0], r["f"][1], r["f"][2]
            ref = ops.pack_dgrad_weight_dev(w, lo, lo + n, r["rows"], cout_pad, eng.dt)
        assert torch.equal(r["out"].view(torch.int16), ref.view(torch.int16)), key


@pytest.mark.gpu
def test_copy_multi_gathers_strided_and_contiguous_vectors(cuda_lib):
    """ops.CopyMulti (prn_copy_multi_f32): contiguous sources of odd lengths / unaligned offsets and strided column views into
    slices of one flat buffer, bit for bit, with and without a scale; tables filled after the launch was recorded."""
    from planerecnet_b200 import ops
    g = torch.Generator().manual_seed(0)
    srcs = [torch.randn(n, generator=g).cuda() for n in (1, 7, 2048, 2049, 5000, 12345)]
    acc = torch.randn(300, 2, generator=g).cuda()
    srcs += [acc[:257, 0], acc[:, 1], torch.randn(64, 3, 3, 3, generator=g).cuda(), torch.randn(9, 4, generator=g).cuda().t()]
    total = sum(s.numel() for s in srcs)
    flat = torch.full((total + 3,), -7.0, device="cuda")
    dsts, off = [], 3                                     # an odd offset: unaligned destinations
    for s in srcs:
        dsts.append(flat[off:off + s.numel()].view(s.shape))
        off += s.numel()
    cm = ops.CopyMulti(dsts)
    norm = cm.normalise(srcs)
    cm.set_sources(norm)
    cm.run()
    want = torch.cat([s.reshape(-1) for s in srcs])
    assert torch.equal(flat[3:], want) and float(flat[0]) == -7.0
    cm.run(scale=0.5)
    assert torch.equal(flat[3:], want * 0.5)


@pytest.mark.gpu
def test_flat_gradient_buffer_of_the_graphed_step_and_autograd_handover(cuda_lib):
    """GraphedStep(flat_grads=True): the backward graph's last launch gathers every parameter gradient into one flat buffer whose
    views equal the per-parameter gradients bit for bit; the autograd boundary (use_train_graph) hands out views of ONE clone of
    it: .grad of every parameter equals the graph's gradient and does not alias the step's static buffer."""
    from planerecnet_b200.train_engine import GraphedStep
    from planerecnet_b200.utils.synth import make_cotangents
    net = TC._build("PlaneRecNet_50_config", cond=True).train().cuda()
    x = torch.randn(2, 3, 128, 160, generator=torch.Generator().manual_seed(5)).cuda()
    eng = net.train_engine
    step = GraphedStep(eng, net, x, flat_grads=True)
    cots = make_cotangents(step.outs, seed=3, device="cuda")
    step.forward(x)
    grads = step.backward(*cots)
    torch.cuda.synchronize()
    params = [p for p in net.parameters() if p.requires_grad]
    assert step.flat is not None and step.flat.numel() == sum(p.numel() for p in params)
    for p in params:
        assert torch.equal(step.flat_views[id(p)], grads[id(p)].reshape(p.shape).float()), p.shape
    # autograd boundary
    net.use_train_graph = True
    outs = net(x)
    loss = (outs[0] * cots[0]).sum() + sum((c * d).sum() for c, d in zip(outs[1], cots[1])) + \
        sum((k * d).sum() for k, d in zip(outs[2], cots[2])) + (outs[3] * cots[3]).sum()
    loss.backward()
    torch.cuda.synchronize()
    st = next(iter(eng._graphs_t.values()))
    lo, hi = st.flat.data_ptr(), st.flat.data_ptr() + st.flat.numel() * 4
    n_checked = 0
    for p in params:
        assert p.grad is not None and bool(torch.isfinite(p.grad).all())
        assert not (lo <= p.grad.data_ptr() < hi)
        assert torch.equal(p.grad, st.flat_views[id(p)]), p.shape
        n_checked += 1
    assert n_checked == len(params)


@pytest.mark.gpu
def test_fused_adam_skips_a_step_with_nonfinite_gradients(cuda_lib):
    """FusedAdam(skip_nonfinite=True) (prn_adam_multi_checked): a step whose gradients contain one inf (or NaN) changes neither the
    parameters, nor the moments, nor the step counter and raises found_inf; the next clean step equals torch.optim.Adam's first."""
    from planerecnet_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(1)
    shapes = [(70000,), (33, 5, 3, 3), (9,)]
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    our_p = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    ref = torch.optim.Adam(ref_p, lr=1e-2)
    ours = FusedAdam(our_p, lr=1e-2, skip_nonfinite=True)
    grads = [torch.randn(s, generator=g).cuda() for s in shapes]
    for bad in (float("inf"), float("nan")):
        poisoned = [gr.clone() for gr in grads]
        poisoned[0][65536 + 11] = bad                       # in the second chunk of the first tensor
        before = [p.detach().clone() for p in our_p]
        ours.step({id(p): gr for p, gr in zip(our_p, poisoned)})
        torch.cuda.synchronize()
        assert int(ours.found_inf) == 1 and float(ours._state3[0]) == 0.0
        for p, b in zip(our_p, before):
            assert torch.equal(p.detach(), b)
        for p in our_p:
            assert float(ours.state[id(p)]["exp_avg"].abs().max()) == 0.0
    for p, gr in zip(ref_p, grads):
        p.grad = gr.clone()
    ref.step()
    ours.step({id(p): gr for p, gr in zip(our_p, grads)})
    torch.cuda.synchronize()
    assert int(ours.found_inf) == 0 and float(ours._state3[0]) == 1.0
    for a, b in zip(our_p, ref_p):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6), float((a - b).abs().max())
