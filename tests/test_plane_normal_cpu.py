"""The batched plane surface-normal term (planerecnet_b200.losses._PlaneNormalBatched: all planes of all images in a few dozen
tensor ops, one global sort for the per-region worst-75 % tails) against the per-plane formulation that mirrors
models/functions/vnl.py:6-165 line by line (losses._PlaneNormal, itself pinned to the unmodified reference through
tests/golden/loss_golden.pt): identical numpy-RNG triplets, values and gradients on the CPU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import loss_cases as LC
from planerecnet_b200 import losses as PL


def _both(name, sampling="numpy"):
    _, _, _, depth, gts, gt_depth = LC.synth(**LC.CASES[name])
    B = depth.shape[0]
    d0 = depth.clone().requires_grad_(True)
    up0 = F.interpolate(d0, scale_factor=2, mode="bilinear", align_corners=False)
    ref_fn = PL._PlaneNormal((480, 640))
    np.random.seed(0)
    ref = torch.stack([ref_fn(up0[b], gts[b]["masks"].bool(), gts[b]["plane_paras"][:, :3], gt_depth[b], gts[b]["k_matrix"])
                       for b in range(B)])
    d1 = depth.clone().requires_grad_(True)
    up1 = F.interpolate(d1, scale_factor=2, mode="bilinear", align_corners=False)
    np.random.seed(0)
    got = PL._PlaneNormalBatched((480, 640), sampling=sampling)(up1, gts, gt_depth)
    return ref, got, d0, d1


@pytest.mark.parametrize("name", list(LC.CASES))
def test_batched_plane_normal_term_equals_the_per_plane_loop(name):
    ref, got, d0, d1 = _both(name)
    assert got.shape == ref.shape
    nan_r, nan_g = torch.isnan(ref), torch.isnan(got)
    assert torch.equal(nan_r, nan_g), (ref, got)                       # the reference's NaN for a degenerate plane is kept
    ok = ~nan_r
    assert torch.allclose(got[ok].double(), ref[ok].double(), rtol=1e-6, atol=1e-8), (ref, got)
    # gradients of the finite images (the weights of a NaN image are NaN in both formulations)
    torch.nansum(ref).backward()
    torch.nansum(got).backward()
    if bool(ok.any()):
        a, b = d1.grad[ok], d0.grad[ok]
        assert float((a - b).abs().max()) <= 1e-6 * float(b.abs().max() + 1e-12), float((a - b).abs().max())
        assert float(b.abs().max()) > 0


def test_device_sampling_is_statistically_equivalent():
    """sampling='device' draws the same distribution of triplets from torch's RNG: the loss agrees to sampling noise."""
    ref, _, _, _ = _both("loss_seed0")
    torch.manual_seed(0)
    vals = []
    for _ in range(3):
        _, got, _, _ = _both("loss_seed0", sampling="device")
        vals.append(got)
    got = torch.stack(vals).mean(0)
    assert torch.allclose(got.double(), ref.double(), rtol=5e-2), (ref, got)


@pytest.mark.parametrize("seed", [0, 1, 7, 12345])
def test_c_restatement_of_numpy_choice_shuffle_is_bit_identical(seed):
    """prn_numpy_choice_shuffle (host C, csrc/prn_hostrng.cu) against the literal np.random.choice + np.random.shuffle calls of
    vnl.py:48-53: same indices for every region and repeat, and the same global RNG stream afterwards (a pure host function of
    the C-ABI library: no device involved)."""
    rs = np.random.RandomState(seed)
    counts = [int(c) for c in rs.randint(0, 120000, size=23)] + [0, 1, 2, 3, 4, 7, 307200, 65536, 65537]
    is_rest = [bool(i % 5 == 4) for i in range(len(counts))]
    fn = PL._PlaneNormalBatched((480, 640), sampling="numpy")
    T = sum(int(c * 0.3) for c in counts)
    a = np.full((3, T + 5), -1, dtype=np.int32)
    b = np.full((3, T + 5), -1, dtype=np.int32)
    np.random.seed(seed)
    np.random.random_sample(seed % 613)                      # move the stream position off a block boundary
    ks_a = fn._sample_host_numpy(counts, is_rest, a)
    after_a = np.random.get_state()
    tail_a = np.random.randint(0, 1 << 30, size=8)
    np.random.seed(seed)
    np.random.random_sample(seed % 613)
    ks_b = fn._sample_host(counts, is_rest, b)
    after_b = np.random.get_state()
    tail_b = np.random.randint(0, 1 << 30, size=8)
    assert ks_a == ks_b and sum(ks_a) > 0
    assert np.array_equal(a, b)
    assert after_a[2] == after_b[2] and np.array_equal(after_a[1], after_b[1])
    assert np.array_equal(tail_a, tail_b)
    for k, c, rest in zip(ks_a, counts, is_rest):
        assert k == (0 if (rest and c == 0) else int(c * 0.3))


def test_numpy_py_sampling_gives_the_same_loss_as_the_c_sampler():
    _, got_c, _, _ = _both("loss_seed0", sampling="numpy")
    _, got_py, _, _ = _both("loss_seed0", sampling="numpy_py")
    assert torch.equal(torch.nan_to_num(got_c), torch.nan_to_num(got_py))


def test_c_sampler_property_random_region_lists():
    """Property test (hypothesis) of prn_numpy_choice_shuffle: arbitrary region sizes (incl. powers of two and their neighbours,
    where the rejection mask changes), arbitrary draw counts k <= n, arbitrary stream positions — indices and the stream afterwards
    equal numpy's for every example."""
    import ctypes as C
    from hypothesis import given, settings, strategies as st
    from planerecnet_b200 import _lib as L

    sizes = st.one_of(st.integers(1, 5000), st.sampled_from([1, 2, 3, 4, 5, 7, 8, 9, 255, 256, 257, 1023, 1024, 1025, 4095, 4096, 4097]))

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.tuples(sizes, st.floats(0.0, 1.0)), min_size=1, max_size=6), st.integers(0, 2 ** 31 - 1), st.integers(0, 700),
           st.integers(1, 3))
    def check(regions, seed, burn, repeats):
        ns = [n for n, _ in regions]
        ks = [int(n * f) for n, f in regions]
        T = sum(ks)
        np.random.seed(seed)
        np.random.random_sample(burn)
        ref = np.full((repeats, T + 1), -1, dtype=np.int32)
        off = 0
        for n, k in zip(ns, ks):
            if k:
                for j in range(repeats):
                    p = np.random.choice(n, k, replace=True)
                    np.random.shuffle(p)
                    ref[j, off:off + k] = p
            off += k
        ref_state = np.random.get_state()
        np.random.seed(seed)
        np.random.random_sample(burn)
        s0 = np.random.get_state()
        key = np.ascontiguousarray(s0[1], dtype=np.uint32).copy()
        pos = C.c_int32(int(s0[2]))
        out = np.full((repeats, T + 1), -1, dtype=np.int32)
        n_arr, k_arr = np.asarray(ns, dtype=np.int64), np.asarray(ks, dtype=np.int64)
        L.check(L.lib().prn_numpy_choice_shuffle(key.ctypes.data_as(C.c_void_p), C.byref(pos), n_arr.ctypes.data_as(C.c_void_p),
                                                 k_arr.ctypes.data_as(C.c_void_p), len(ns), repeats, out.ctypes.data_as(C.c_void_p),
                                                 out.strides[0] // 4), "prn_numpy_choice_shuffle")
        assert np.array_equal(out, ref)
        assert pos.value == ref_state[2] and np.array_equal(key, ref_state[1])

    check()
