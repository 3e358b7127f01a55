"""Thin Python wrappers over the C ABI: build the plain-C descriptors from torch tensors (torch is only
the owner of device memory and streams here) and launch on the current CUDA stream."""
import ctypes as C

import torch

from . import _lib as L


def torch_dtype(dtype):
    return torch.bfloat16 if dtype == L.PRN_BF16 else torch.float16


def round_up(x, m):
    return (x + m - 1) // m * m


def pack_conv_weight(w, c_splits=None, n_pad=None, dtype=L.PRN_BF16, scale=None):
    """[Cout, Cin, kh, kw] fp32 -> packed [n_pad, kh*kw*sum(c_pad)] 16-bit, K ordered (ky, kx, c).

    c_splits: list of (real_channels, padded_channels) per concatenated source; channels of each source
    are zero-padded to its padded width.  scale: optional per-output-channel fp32 factor (folded BN)."""
    cout, cin, kh, kw = w.shape
    w = w.detach().float()
    if scale is not None:
        w = w * scale.view(-1, 1, 1, 1).float()
    if c_splits is None:
        c_splits = [(cin, round_up(cin, 64))]
    assert sum(r for r, _ in c_splits) == cin
    parts, off = [], 0
    for real, padded in c_splits:
        blk = w[:, off:off + real]
        if padded > real:
            blk = torch.nn.functional.pad(blk, (0, 0, 0, 0, 0, padded - real))
        parts.append(blk)
        off += real
    w = torch.cat(parts, dim=1)                       # [Cout, Cpad, kh, kw]
    w = w.permute(0, 2, 3, 1).reshape(cout, -1)       # [Cout, kh*kw*Cpad]
    n_pad = n_pad or round_up(cout, 16)
    if n_pad > cout:
        w = torch.nn.functional.pad(w, (0, 0, 0, n_pad - cout))
    return w.to(torch_dtype(dtype)).contiguous()


def subpixel_weights(w):
    """[Cout, Cin, 3, 3] -> [4*Cout, Cin, 3, 3]: the four sub-pixel phases of Upsample(x2, nearest) -> ReflectionPad2d(1) ->
    Conv2d(3x3) (planerecnet.py:540-567) as one 3x3 convolution at the LOW resolution with replicate padding.  Output pixel
    (2y+a, 2x+b) reads upsampled rows 2y+a-1 .. 2y+a+1 = low-resolution rows {y-1, y, y} (a = 0) or {y, y, y+1} (a = 1); the
    reflected border row of the upsampled map is the clamped low-resolution row.  Phase (a, b) occupies output rows
    [(2a+b)*Cout, +Cout); taps that fall on the same low-resolution pixel are summed (in fp32, before the 16-bit rounding)."""
    cout, cin, kh, kw = w.shape
    assert kh == 3 and kw == 3
    w = w.detach().float()
    sets = {0: ([0], [1, 2], []), 1: ([], [0, 1], [2])}      # phase -> source taps of low-resolution tap 0, 1, 2
    out = torch.zeros(4, cout, cin, 3, 3, dtype=torch.float32, device=w.device)
    for a in (0, 1):
        for b in (0, 1):
            for ty in range(3):
                for tx in range(3):
                    for ky in sets[a][ty]:
                        for kx in sets[b][tx]:
                            out[2 * a + b, :, :, ty, tx] += w[:, :, ky, kx]
    return out.reshape(4 * cout, cin, 3, 3)


def pad_vec(v, n_pad):
    v = v.detach().float()
    if v.numel() < n_pad:
        v = torch.nn.functional.pad(v, (0, n_pad - v.numel()))
    return v.contiguous()


def conv2d(src0, weight, *, batch, h_in, w_in, ksize=1, stride=1, pad=0, pad_mode=L.PAD_ZERO, upsample=1,
           src1=None, bias=None, residual=None, act=L.ACT_NONE, act_param=0.0, out16=None, out32=None,
           ld_out16=None, ld_out32=None, out_img_rows=0, stats=None, stats_cg=0, dcn_offmask=None,
           n_pad=None, w_group_rows=0, dtype=L.PRN_BF16, c0=None, c1=None, ld_res=None, counters=None, shuffle_n=0):
    """Launch prn_conv2d_fwd.  src*/residual/out16 are 16-bit NHWC tensors (any shape, channels last)."""
    d = L.PrnConv()
    d.src0 = src0.data_ptr()
    d.c0 = c0 if c0 is not None else src0.shape[-1]
    d.src1 = src1.data_ptr() if src1 is not None else None
    d.c1 = (c1 if c1 is not None else src1.shape[-1]) if src1 is not None else 0
    d.ld0 = src0.shape[-1] if src0.shape[-1] != d.c0 else 0
    d.ld1 = (src1.shape[-1] if src1.shape[-1] != d.c1 else 0) if src1 is not None else 0
    d.batch, d.h_in, d.w_in = batch, h_in, w_in
    d.upsample = upsample
    d.ksize, d.stride, d.pad, d.pad_mode = ksize, stride, pad, pad_mode
    d.h_out = (h_in * upsample + 2 * pad - ksize) // stride + 1
    d.w_out = (w_in * upsample + 2 * pad - ksize) // stride + 1
    d.dcn_offmask = dcn_offmask.data_ptr() if dcn_offmask is not None else None
    d.weight = weight.data_ptr()
    d.w_rows_total = weight.shape[0]
    d.n_pad = n_pad if n_pad is not None else weight.shape[0]
    d.w_group_rows = w_group_rows
    d.bias = bias.data_ptr() if bias is not None else None
    d.residual = residual.data_ptr() if residual is not None else None
    d.ld_res = (ld_res if ld_res is not None else residual.shape[-1]) if residual is not None else 0
    d.act, d.act_param = act, act_param
    d.out16 = out16.data_ptr() if out16 is not None else None
    d.ld_out16 = (ld_out16 if ld_out16 is not None else out16.shape[-1]) if out16 is not None else 0
    d.out32 = out32.data_ptr() if out32 is not None else None
    d.ld_out32 = (ld_out32 if ld_out32 is not None else out32.shape[-1]) if out32 is not None else 0
    d.out_img_rows = out_img_rows
    d.stats = stats.data_ptr() if stats is not None else None
    d.stats_cg = stats_cg
    d.dtype = dtype
    d.shuffle_n = shuffle_n
    if counters is not None:
        L.check(L.lib().prn_conv2d_fwd_profile(C.byref(d), L.current_stream(), C.c_void_p(counters.data_ptr())),
                "prn_conv2d_fwd_profile")
    else:
        L.check(L.lib().prn_conv2d_fwd(C.byref(d), L.current_stream()), "prn_conv2d_fwd")
    return d


# ------------------------------------------------------------------------------------------ training step
def _vp(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def pack_dgrad_weight(w, cout_pad=None, n_pad=None, dtype=L.PRN_BF16):
    """[Cout, Cin, kh, kw] -> packed weights of the input-gradient convolution: rows = Cin (padded to n_pad),
    K = (ky, kx, Cout padded), taps flipped.  dX = conv(dY, this, pad = k - 1 - pad) for stride 1."""
    wt = w.detach().float().flip(2, 3).transpose(0, 1).contiguous()      # [Cin, Cout, kh, kw]
    cout = w.shape[0]
    return pack_conv_weight(wt, [(cout, cout_pad or round_up(cout, 64))], n_pad or round_up(w.shape[1], 16), dtype)


def conv2d_wgrad(src0, dy, dw, *, batch, h_in, w_in, n, ksize=1, stride=1, pad=0, pad_mode=L.PAD_ZERO, upsample=1,
                 src1=None, c0=None, c1=None, dtype=L.PRN_BF16, flags=0):
    """Launch prn_conv2d_wgrad: dw[n, ksize*ksize*(c0+c1)] (fp32, accumulated) += dy^T . im2col(src)."""
    d = L.PrnWgrad()
    d.src0 = src0.data_ptr()
    d.c0 = c0 if c0 is not None else src0.shape[-1]
    d.src1 = src1.data_ptr() if src1 is not None else None
    d.c1 = (c1 if c1 is not None else src1.shape[-1]) if src1 is not None else 0
    d.ld0 = src0.shape[-1] if src0.shape[-1] != d.c0 else 0
    d.ld1 = (src1.shape[-1] if src1.shape[-1] != d.c1 else 0) if src1 is not None else 0
    d.batch, d.h_in, d.w_in = batch, h_in, w_in
    d.upsample = upsample
    d.ksize, d.stride, d.pad, d.pad_mode = ksize, stride, pad, pad_mode
    d.h_out = (h_in * upsample + 2 * pad - ksize) // stride + 1
    d.w_out = (w_in * upsample + 2 * pad - ksize) // stride + 1
    d.dy = dy.data_ptr()
    d.n = n
    d.ld_dy = dy.shape[-1]
    d.dw = dw.data_ptr()
    d.ld_dw = dw.shape[-1]
    d.dtype = dtype
    d.flags = flags
    L.check(L.lib().prn_conv2d_wgrad(C.byref(d), L.current_stream()), "prn_conv2d_wgrad")
    return d


def unpack_wgrad(dw, weight_shape, c_splits=None):
    """fp32 [n_rows, kh*kw*sum(c_pad)] (layout of pack_conv_weight) -> [Cout, Cin, kh, kw] fp32."""
    cout, cin, kh, kw = weight_shape
    if c_splits is None:
        c_splits = [(cin, round_up(cin, 64))]
    cpad = sum(p for _, p in c_splits)
    v = dw[:cout].view(cout, kh, kw, cpad)
    if len(c_splits) == 1:                            # one source: a single strided copy (torch.cat of one part is a copy of its own)
        return v[..., :c_splits[0][0]].permute(0, 3, 1, 2).contiguous()
    parts, off = [], 0
    for real, padded in c_splits:
        parts.append(v[..., off:off + real])
        off += padded
    return torch.cat(parts, dim=-1).permute(0, 3, 1, 2).contiguous()


def bn_finalize(stats, mean_invstd, running_mean, running_var, count, eps, momentum):
    L.check(L.lib().prn_bn_finalize(_vp(stats), _vp(mean_invstd), _vp(running_mean), _vp(running_var), stats.numel() // 2,
                                    C.c_int64(count), C.c_float(eps), C.c_float(momentum), L.current_stream()), "prn_bn_finalize")


def bn_apply(x, out, mean_invstd, gamma, beta, residual, relu, dtype):
    c = x.shape[-1]
    L.check(L.lib().prn_bn_apply(_vp(x), _vp(out), _vp(mean_invstd), _vp(gamma), _vp(beta), _vp(residual),
                                 C.c_int64(x.numel() // c), c, 1 if relu else 0, dtype, L.current_stream()), "prn_bn_apply")


def bn_finalize_apply(x, out, stats, mean_invstd, running_mean, running_var, count, eps, momentum, gamma, beta, residual, relu, dtype):
    c = x.shape[-1]
    L.check(L.lib().prn_bn_finalize_apply(_vp(x), _vp(out), _vp(stats), _vp(mean_invstd), _vp(running_mean), _vp(running_var),
                                          C.c_int64(count), C.c_float(eps), C.c_float(momentum), _vp(gamma), _vp(beta), _vp(residual),
                                          C.c_int64(x.numel() // c), c, 1 if relu else 0, dtype, L.current_stream()), "prn_bn_finalize_apply")


def chan_reduce(dz, out, x, mean_invstd, sums, dtype):
    c = dz.shape[-1]
    L.check(L.lib().prn_chan_reduce(_vp(dz), _vp(out), _vp(x), _vp(mean_invstd), _vp(sums), C.c_int64(dz.numel() // c), c, dtype,
                                    L.current_stream()), "prn_chan_reduce")


def bn_bwd_apply(dz, out, x, mean_invstd, gamma, sums, dx, g_out, dtype):
    c = dz.shape[-1]
    L.check(L.lib().prn_bn_bwd_apply(_vp(dz), _vp(out), _vp(x), _vp(mean_invstd), _vp(gamma), _vp(sums), _vp(dx), _vp(g_out),
                                     C.c_int64(dz.numel() // c), c, dtype, L.current_stream()), "prn_bn_bwd_apply")


def relu_bwd(dz, out, g, dtype):
    L.check(L.lib().prn_relu_bwd(_vp(dz), _vp(out), _vp(g), C.c_int64(dz.numel()), dtype, L.current_stream()), "prn_relu_bwd")


def add_strided(dst, src, stride, dtype):
    b, h, w, c = src.shape
    L.check(L.lib().prn_add_strided(_vp(dst), _vp(src), b, h, w, c, stride, dtype, L.current_stream()), "prn_add_strided")


def add_f32(a32, b16, out16, dtype):
    L.check(L.lib().prn_add_f32(_vp(a32), _vp(b16), _vp(out16), C.c_int64(a32.numel()), dtype, L.current_stream()), "prn_add_f32")


def add16(a, b, out, dtype):
    L.check(L.lib().prn_add16(_vp(a), _vp(b), _vp(out), C.c_int64(a.numel()), dtype, L.current_stream()), "prn_add16")


def maxpool_bwd(x, dout, din, dtype):
    b, h, w, c = x.shape
    L.check(L.lib().prn_maxpool3x3s2_bwd(_vp(x), _vp(dout), _vp(din), b, h, w, c, dtype, L.current_stream()), "prn_maxpool3x3s2_bwd")


def dcn_im2col(x, offmask, col, stride, pad, dtype):
    b, h, w, c = x.shape
    L.check(L.lib().prn_dcn_im2col(_vp(x), _vp(offmask), _vp(col), b, h, w, c, stride, pad, dtype, L.current_stream()), "prn_dcn_im2col")


def dcn_col2im_bwd(x, offmask, dcol, dx32, dpre16, stride, pad, bound, dtype):
    b, h, w, c = x.shape
    L.check(L.lib().prn_dcn_col2im_bwd(_vp(x), _vp(offmask), _vp(dcol), _vp(dx32), _vp(dpre16), b, h, w, c, stride, pad,
                                       C.c_float(bound), dtype, L.current_stream()), "prn_dcn_col2im_bwd")


def gn_bwd_reduce(dz, out, x, stats, sums_bc, dgb, cg, eps, dtype):
    b, c = dz.shape[0], dz.shape[-1]
    hw = dz.numel() // (b * c)
    L.check(L.lib().prn_gn_bwd_reduce(_vp(dz), _vp(out), _vp(x), _vp(stats), _vp(sums_bc), _vp(dgb), b, hw, c, cg, C.c_float(eps),
                                      dtype, L.current_stream()), "prn_gn_bwd_reduce")


def gn_bwd_apply(dz, out, x, stats, gamma, sums_bc, dx, cg, eps, dtype):
    b, c = dz.shape[0], dz.shape[-1]
    hw = dz.numel() // (b * c)
    L.check(L.lib().prn_gn_bwd_apply(_vp(dz), _vp(out), _vp(x), _vp(stats), _vp(gamma), _vp(sums_bc), _vp(dx), b, hw, c, cg,
                                     C.c_float(eps), dtype, L.current_stream()), "prn_gn_bwd_apply")


def avgpool2_bwd(dout, din, accumulate, dtype):
    b, h, w, c = din.shape
    L.check(L.lib().prn_avgpool2x2_bwd(_vp(dout), _vp(din), b, h, w, c, 1 if accumulate else 0, dtype, L.current_stream()),
            "prn_avgpool2x2_bwd")


def upsample2x_bwd(dout, din, dtype):
    b, h, w, c = din.shape
    L.check(L.lib().prn_upsample2x_bilinear_bwd(_vp(dout), _vp(din), b, h, w, c, dtype, L.current_stream()),
            "prn_upsample2x_bilinear_bwd")


def resize_bilinear_bwd(dout, din32, dtype):
    b, h, w, c = din32.shape
    L.check(L.lib().prn_resize_bilinear_bwd(_vp(dout), _vp(din32), b, h, w, c, dout.shape[1], dout.shape[2], dout.shape[3], dtype,
                                            L.current_stream()), "prn_resize_bilinear_bwd")


def reflect_fold(dpad, din, upsample, accumulate, dtype):
    b, h, w, c = din.shape
    L.check(L.lib().prn_reflect_fold(_vp(dpad), _vp(din), b, h, w, c, dpad.shape[-1], upsample, 1 if accumulate else 0, dtype,
                                     L.current_stream()), "prn_reflect_fold")


def softplus_bwd_pad(dout32, out32, dpre16, dtype):
    L.check(L.lib().prn_softplus_bwd_pad(_vp(dout32), _vp(out32), _vp(dpre16), C.c_int64(out32.numel()), dtype, L.current_stream()),
            "prn_softplus_bwd_pad")


def pack_conv_weight_dev(w, c_splits, n_pad, dtype):
    """prn_pack_conv_weight: same result as pack_conv_weight, one launch, for a CUDA fp32 [Cout,Cin,k,k] parameter."""
    cout, cin, k, _ = w.shape
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and 1 <= len(c_splits) <= 2
    los, off = [], 0
    for real, _ in c_splits:
        los.append(off)
        off += real
    assert off == cin
    arr = C.c_int32 * len(c_splits)
    out = torch.empty(n_pad, k * k * sum(p for _, p in c_splits), dtype=torch_dtype(dtype), device="cuda")
    L.check(L.lib().prn_pack_conv_weight(_vp(w), _vp(out), cout, cin, k, n_pad, len(c_splits), arr(*los),
                                         arr(*[r for r, _ in c_splits]), arr(*[p for _, p in c_splits]), dtype,
                                         L.current_stream()), "prn_pack_conv_weight")
    return out


def pack_dgrad_weight_dev(w, lo, hi, rows_pad, cout_pad, dtype):
    """prn_pack_dgrad_weight: == pack_dgrad_weight(w[:, lo:hi], cout_pad, rows_pad), one launch."""
    cout, cin, k, _ = w.shape
    assert w.is_cuda and w.dtype == torch.float32 and w.is_contiguous()
    out = torch.empty(rows_pad, k * k * cout_pad, dtype=torch_dtype(dtype), device="cuda")
    L.check(L.lib().prn_pack_dgrad_weight(_vp(w), _vp(out), cout, cin, k, lo, hi, rows_pad, cout_pad, dtype, L.current_stream()),
            "prn_pack_dgrad_weight")
    return out


class CopyMulti:
    """Many fp32 vectors -> contiguous destinations with ONE launch (prn_copy_multi_f32): the parameter gradients of a training
    step gathered into the flat buffer the data-parallel all-reduce (and autograd's hand-over) works on.  torch._foreach_copy_
    falls back to one copy per tensor as soon as one source is strided (bias / BatchNorm gradients are column views of their
    accumulators): ~400 launches per step.

    For CUDA graphs: construct with the destinations (allocates the device tables), call normalise() + run() inside the capture
    (the launch only references the tables), and set_sources() after the capture has ended (uploads the table contents, which the
    kernel reads at replay time; source addresses are only known once the captured backward has allocated them)."""
    CHUNK = 2048

    def __init__(self, dsts):
        self.dsts = list(dsts)
        work = []
        for i, d in enumerate(self.dsts):
            assert d.dtype == torch.float32 and d.is_contiguous()
            work += [(i, c) for c in range((d.numel() + self.CHUNK - 1) // self.CHUNK)]
        self.recs = torch.zeros(32 * len(self.dsts), dtype=torch.uint8, device="cuda")
        self.work = torch.tensor(work, dtype=torch.int32).cuda().contiguous()
        self.n_blocks = len(work)
        self.srcs = None

    @staticmethod
    def normalise(srcs):
        """Sources the kernel can read: contiguous, or 1-D with a stride; anything else is made contiguous (one copy)."""
        return [s if (s.is_contiguous() or s.dim() == 1) else s.contiguous() for s in srcs]

    def set_sources(self, srcs):
        import struct
        assert not torch.cuda.is_current_stream_capturing(), "set_sources uploads the tables: call it outside the capture"
        assert len(srcs) == len(self.dsts)
        blob = b""
        for d, s in zip(self.dsts, srcs):
            assert s.dtype == torch.float32 and s.numel() == d.numel() and (s.is_contiguous() or s.dim() == 1)
            blob += struct.pack("<QQqii", s.data_ptr(), d.data_ptr(), d.numel(), 1 if s.is_contiguous() else s.stride(0), 0)
        self.recs.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
        self.srcs = list(srcs)                       # keeps the sources alive for as long as the tables point at them

    def run(self, scale=1.0):
        L.check(L.lib().prn_copy_multi_f32(_vp(self.recs), _vp(self.work), self.n_blocks, C.c_float(scale), L.current_stream()),
                "prn_copy_multi_f32")


class UnpackMulti:
    """Weight gradients of many convs, accumulator layout -> parameter layout, with ONE launch (prn_unpack_wgrad_multi) instead of a
    permute + contiguous per conv.  Built for CUDA-graph capture: the device tables are allocated up front, add() / run() are
    called while the backward is being captured (pointers and the grid size are known then), flush() uploads the table contents
    after the capture has ended (the launch reads them at replay time)."""
    ROWS = 8

    def __init__(self, max_recs=1024, max_work=1 << 17):
        self.recs = torch.zeros(40 * max_recs, dtype=torch.uint8, device="cuda")
        self.work = torch.zeros(max_work, 2, dtype=torch.int32, device="cuda")
        self.max_recs, self.max_work = max_recs, max_work
        self.items, self.keep, self.ran = [], [], False

    def add(self, dw, dst, weight_shape, cpad):
        cout, cin, kh, kw = weight_shape
        assert dw.dtype == torch.float32 and dst.dtype == torch.float32 and dst.is_contiguous() and dw.stride(1) == 1
        assert dw.shape[0] >= cout and dw.shape[1] >= kh * kw * cpad and cin <= cpad and dst.numel() == cout * cin * kh * kw
        self.items.append((dw.data_ptr(), dst.data_ptr(), cout, cin, kh * kw, cpad, dw.stride(0)))
        self.keep += [dw, dst]

    def run(self):
        assert not self.ran and self.items and len(self.items) <= self.max_recs
        self.n_blocks = sum((it[2] + self.ROWS - 1) // self.ROWS for it in self.items)
        assert self.n_blocks <= self.max_work
        self.smem = max(it[4] * (it[5] + 1) * 4 for it in self.items)
        L.check(L.lib().prn_unpack_wgrad_multi(_vp(self.recs), _vp(self.work), self.n_blocks, self.smem, L.current_stream()),
                "prn_unpack_wgrad_multi")
        self.ran = True

    def flush(self):
        import struct
        assert not torch.cuda.is_current_stream_capturing()
        blob, work = b"", []
        for i, (src, dst, cout, cin, kk, cpad, ld) in enumerate(self.items):
            blob += struct.pack("<QQ6i", src, dst, cout, cin, kk, cpad, ld, 0)          # = sizeof(UnpackRec) = 40
            work += [(i, n) for n in range(0, cout, self.ROWS)]
        self.recs[:len(blob)].copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
        self.work[:len(work)].copy_(torch.tensor(work, dtype=torch.int32))
