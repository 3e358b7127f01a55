"""Parity on the BENCHMARKED configurations (BASELINE.json configs[1..3]: 480x640, batch 8, ResNet50- and ResNet101-DCN), the
shapes bench.py times: M = 153 600 ... 614 400 output rows per launch, > 1 200 tiles per launch, the full-size gather offsets.
Every stage tap of the dense forward against the CPU oracle (pinned to the unmodified reference), the bs = 1 golden fixtures
generated from the unmodified reference at 480x640, and one training step at 480x640 against autograd through the oracle.

Tolerances (rel-L2 per tensor, from BASELINE.json north_star: 1e-2 for 16-bit): f16 3e-3, bf16 2e-2 — as in test_model_gpu.py."""
import functools
import os

import pytest
import torch

import helpers as H
import train_cases as TC
from oracle import prn_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {"f16": 3e-3, "bf16": 2e-2}
H_IMG, W_IMG, B_BENCH = 480, 640, 8


@functools.lru_cache(maxsize=2)
def _oracle_bs8(preset):
    net = H.perturb_(H.build_ours(preset, 0)).eval()
    x = H.make_input(B_BENCH, H_IMG, W_IMG, 0)
    orc = O.Oracle(net.state_dict(), preset)
    with torch.no_grad():
        ref = orc.forward_dense(x)
    return net, x, orc, ref


@pytest.mark.parametrize("preset,prec", [("PlaneRecNet_101_config", "f16"), ("PlaneRecNet_101_config", "bf16"),
                                         ("PlaneRecNet_50_config", "f16")])
def test_dense_forward_bs8_480x640_matches_oracle(cuda_lib, preset, prec):
    net, x, orc, (omask, ocate, okern, odepth) = _oracle_bs8(preset)
    import copy
    netc = copy.deepcopy(net).cuda().set_precision(prec)
    with torch.no_grad():
        st = netc.engine.forward_dense(netc, x.cuda())
        st_g = netc.engine.forward_dense_graph(netc, x.cuda())      # the path bench.py replays
    torch.cuda.synchronize()
    tol = TOL[prec]
    for tag, s in (("eager", st), ("graph", st_g)):
        mask, cate, kern, depth = s["outputs"]
        errs = {"mask": H.rel_l2(mask.cpu(), omask), "depth": H.rel_l2(depth.cpu(), odepth)}
        for i in range(4):
            errs[f"cate{i}"] = H.rel_l2(cate[i].cpu(), ocate[i])
            errs[f"kern{i}"] = H.rel_l2(kern[i].cpu(), okern[i])
            c = s["cs"][i][..., :orc.taps["cs"][i].shape[1]].float().permute(0, 3, 1, 2).cpu()
            errs[f"C{i + 2}"] = H.rel_l2(c, orc.taps["cs"][i])
            errs[f"P{i + 2}"] = H.rel_l2(s["ps"][i].float().permute(0, 3, 1, 2).cpu(), orc.taps["ps"][i])
        errs["attn"] = H.rel_l2(s["attn"].float().permute(0, 3, 1, 2).cpu(), orc.taps["ppa_attn"])
        print(f"{preset}/{prec}/{tag}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
        bad = {k: v for k, v in errs.items() if not v <= tol}
        assert not bad, f"{preset}/{prec}/{tag} bs8 480x640: rel-L2 above {tol}: {bad}"
        assert tuple(mask.shape) == (B_BENCH, 128, 120, 160) and tuple(depth.shape) == (B_BENCH, 1, 240, 320)
    del netc
    torch.cuda.empty_cache()


@pytest.mark.parametrize("fixture", ["r50_b1_480x640", "r101_b1_480x640"])
def test_golden_480x640_from_the_unmodified_reference(cuda_lib, fixture):
    """Samples of the unmodified reference's own stage tensors at 480x640 (tests/golden/make_golden.py), no oracle in between."""
    from golden.make_golden import sample_idx
    g = torch.load(os.path.join(GOLD, fixture + ".pt"))
    net = H.perturb_(H.build_ours(g["preset"], g["seed"])).eval().cuda()
    x = H.make_input(g["B"], g["H"], g["W"], g["seed"]).cuda()
    with torch.no_grad():
        st = net.engine.forward_dense(net, x)
        res = net(x)
    mask, cate, kern, depth = st["outputs"]
    named = [("mask", mask), ("depth", depth)] + [(f"kern{i}", kern[i]) for i in range(4)]
    named += [(f"C{i + 2}", st["cs"][i][..., :g["stages"][f"C{i + 2}"]["shape"][1]].float().permute(0, 3, 1, 2)) for i in range(4)]
    named += [(f"P{i + 2}", st["ps"][i].float().permute(0, 3, 1, 2)) for i in range(4)]
    named += [("attn", st["attn"].float().permute(0, 3, 1, 2))]
    for name, t in named:
        gs = g["stages"][name]
        assert list(t.shape) == gs["shape"], (name, t.shape, gs["shape"])
        flat = t.float().contiguous().cpu().flatten()
        got = flat[sample_idx(flat.numel())]
        err = float((got - gs["samples"]).norm() / gs["samples"].norm())
        assert err <= TOL["f16"], f"{fixture} {name}: {err:.3g}"
    for i in range(4):
        assert H.rel_l2(cate[i].cpu(), g["cate_full"][i]) <= TOL["f16"]
    assert H.rel_l2(depth.cpu()[:, :, ::4, ::4], g["depth_ds4"]) <= TOL["f16"]
    # detections of the reference on the same input: count within the threshold-flip margin, best score close
    r, o = res[0], g["results"][0]
    n_r = 0 if r["pred_scores"] is None else len(r["pred_scores"])
    n_o = 0 if o["pred_scores"] is None else len(o["pred_scores"])
    assert abs(n_r - n_o) <= max(2, n_o // 10), (n_r, n_o)
    if n_r and n_o:
        assert abs(float(r["pred_scores"].max()) - float(o["pred_scores"].max())) < 5e-3


@pytest.mark.parametrize("preset", ["PlaneRecNet_101_config", "PlaneRecNet_50_config"])
def test_train_step_480x640_frozen_bn(cuda_lib, preset):
    """One training step (fwd + bwd through the sm_100a tape) at the benchmarked resolution, bs 2, frozen BatchNorm statistics
    (stable network: pins the whole tape and every gradient join), against autograd through the CPU oracle."""
    r = TC.run_model_check(preset, B=2, H=H_IMG, W=W_IMG, prec="bf16", bn_mode="frozen", cond=False)
    print(preset, "outs", [f"{v:.1e}" for v in r["outs"]], "cos", f"{r['all_cos']:.5f}", "rel", f"{r['all_rel']:.2e}")
    assert not r["missing"], r["missing"]
    assert max(r["outs"]) <= TOL["bf16"], r["outs"]
    assert r["all_cos"] >= 0.9995, r["all_cos"]
    torch.cuda.empty_cache()


def test_train_step_480x640_bs8_graph_equals_eager(cuda_lib):
    """The graphed training step bench.py times (R101, bs 8, 480x640, batch-statistics BatchNorm): replayed outputs and
    gradients equal the eager launch sequence up to run-to-run noise (fp32 atomics), every parameter has a finite gradient."""
    from planerecnet_b200.train_engine import GraphedStep
    from planerecnet_b200.utils.synth import make_cotangents
    # residual-dominant init (bn3.weight x0.1, like a trained ResNet): train-mode BatchNorm on the plain random init is chaotic
    # in the fp32 oracle itself, which would make any two runs of the same step differ by tens of percent
    net = TC._build("PlaneRecNet_101_config", cond=True).train().cuda()
    x = H.make_input(B_BENCH, H_IMG, W_IMG, 0).cuda()
    eng = net.train_engine

    def flat(g):
        return torch.cat([g[id(p)].flatten().float() for p in net.parameters() if id(p) in g]).clone()

    def cosine(a, b):
        return float((a.double() @ b.double()) / (a.double().norm() * b.double().norm()))

    def eager():
        outs = eng.forward_train(net, x)
        cots_ = make_cotangents(outs, seed=1, device="cuda")
        eng.seed_output_grads(*cots_)
        g = flat(eng.backward())
        return [outs[0].clone(), outs[3].clone()], g, cots_

    o_e0, g_e0, cots = eager()
    o_e1, g_e1, _ = eager()
    # run-to-run noise of the eager path itself (fp32 atomics in the BatchNorm statistics -> 16-bit rounding flips, amplified by
    # the batch-statistics network): the graph replay has to agree with eager to that noise, not better
    noise_m, noise_d, noise_cos = H.rel_l2(o_e1[0], o_e0[0]), H.rel_l2(o_e1[1], o_e0[1]), cosine(g_e1, g_e0)
    step = GraphedStep(eng, net, x)
    so = step.forward(x)
    gg = step.backward(*cots)
    torch.cuda.synchronize()
    n_params = sum(1 for p in net.parameters() if p.requires_grad)
    assert len(gg) >= n_params, (len(gg), n_params)
    g_g = flat(gg)
    assert g_g.numel() == g_e0.numel() and bool(torch.isfinite(g_g).all())
    e_m, e_d, cos = H.rel_l2(so[0], o_e0[0]), H.rel_l2(so[3], o_e0[1]), cosine(g_g, g_e0)
    print(f"bs8 480x640: eager run-to-run mask {noise_m:.2e} depth {noise_d:.2e} grad cosine {noise_cos:.4f} | "
          f"graph vs eager mask {e_m:.2e} depth {e_d:.2e} grad cosine {cos:.4f}")
    assert e_m <= max(3 * noise_m, 5e-3) and e_d <= max(3 * noise_d, 5e-3), (e_m, e_d, noise_m, noise_d)
    assert (1 - cos) <= max(3 * (1 - noise_cos), 1e-3), (cos, noise_cos)
    del step
    torch.cuda.empty_cache()
