"""Training-step executor of the PlaneRecNet dense path on libprn_b200.

`forward_train` runs the reference's training branch (planerecnet.py:73-103 with net.train(): batch-statistics
BatchNorm, no weight folding) and records one backward closure per launched group on a tape; `backward` replays the
tape in reverse.  Every arithmetic step is a launch of a hand-written sm_100a kernel through the C ABI:

  * forward contractions: prn_conv2d_fwd (per-channel / per-group sums for the normalisations come from its epilogue)
  * input gradients: prn_conv2d_fwd over dY with flipped, in/out-transposed weights (stride 1); reflection padding
    and nearest x2 are folded afterwards (prn_reflect_fold), the 1x1 stride-2 projection is scattered
    (prn_add_strided), the stride-2 offset/modulator conv goes through a zero-inserted dY
  * weight gradients: prn_conv2d_wgrad (split-K tcgen05, fp32 accumulation across levels that share weights)
  * deformable 3x3: prn_dcn_im2col + 1x1 contraction forward, 1x1 contractions + prn_dcn_col2im_bwd backward
  * normalisations / resamplers / pooling: the prn_*_bwd passes of csrc/prn_train*.cu

torch is used for device memory, streams and parameter (un)packing only.  Activations and gradients are NHWC
16-bit (bf16 by default: gradients need the exponent range), weight gradients fp32.
"""
import contextlib
import ctypes as C
import gc

import torch
from torch import nn

from . import _lib as L
from . import ops
from .engine import Engine, _ver
from .models.dcn import DeformableConv2d


class TrainEngine(Engine):
    def __init__(self, dtype="bf16"):
        super().__init__(dtype)
        self.tape = []
        self.grads = {}        # id(activation tensor or token) -> [gradient tensor, owned?]
        self.wbufs = {}        # packed fp32 weight-gradient accumulators, keyed per (conv, channel split)
        self.pgrads = {}       # id(param) -> fp32 gradient in the parameter's shape
        self.pfinal = []       # closures that turn accumulators into parameter gradients after the tape
        self._keep = []        # keeps tokens / tensors alive while their ids are used as keys
        self.side_wgrad = True
        # Loss scaling for 16-bit gradients (needed with f16, whose range ends at 6e-8): cotangents are multiplied by
        # grad_scale on entry; the autograd boundary (and FusedAdam(grad_scale=...)) divide the parameter gradients again.
        self.grad_scale = 1.0
        self._wg_streams = None
        self._wg_rr = 0
        self._wg_used = set()
        self._arena = None
        self._arena_off = 0
        self._arena_hi = 0
        self._graphs_t = {}
        self._nbt = []
        self._marks, self._pf_done = [], 0
        self._pack_recs, self._pack_tab = {}, None
        self._big, self._big_off, self._big_hi, self._big_need = None, 0, 0, 0
        self._defer = None
        self._fin_single = set()

    # ------------------------------------------------------------------ gradient bookkeeping
    def _take(self, t):
        e = self.grads.pop(id(t), None)
        return None if e is None else e[0]

    def _peek(self, t):
        e = self.grads.get(id(t))
        return None if e is None else e[0]

    def _give(self, t, g, owned=True):
        """Join gradient g into the gradient of activation t (never writes into a tensor it does not own)."""
        e = self.grads.get(id(t))
        if e is None:
            self.grads[id(t)] = [g, owned]
            return
        if e[1]:
            ops.add16(e[0], g, e[0], self.dt)
        else:
            s = torch.empty_like(e[0])
            ops.add16(e[0], g, s, self.dt)
            self.grads[id(t)] = [s, True]
        self.launches += 1

    def _set_grad(self, t, g):
        self.grads[id(t)] = [g, True]

    def _padd(self, p, g):
        """Accumulate a parameter gradient (fp32, parameter-shaped)."""
        if p is None or not p.requires_grad:
            return
        cur = self.pgrads.get(id(p))
        if cur is None:
            self.pgrads[id(p)] = g                      # may be a view of an accumulator: copied only if joined
        else:
            assert self._defer is None or id(p) not in self._defer["dst_ids"], "a deferred weight gradient cannot be joined"
            self.pgrads[id(p)] = cur + g

    def _fin_wgrad(self, p, dw, shape, c_splits):
        """Weight gradient of a conv from its accumulator (layout of the packed operand) to the parameter layout.  While a graphed
        step captures its backward (`self._defer`, set by GraphedStep) single-source convs are only RECORDED: one
        prn_unpack_wgrad_multi launch at the end of the backward writes all of them — straight into the flat gradient buffer when
        there is one — instead of a permute + contiguous per conv."""
        if p is None or not p.requires_grad:
            return
        d = self._defer
        if len(c_splits) == 1 and self.pgrads.get(id(p)) is None:
            self._fin_single.add(id(p))                 # (what a later capture of the same step may defer)
        if d is not None and id(p) in d["ids"] and len(c_splits) == 1 and self.pgrads.get(id(p)) is None:
            views = d["views"]
            dst = views[id(p)] if (views is not None and id(p) in views) else torch.empty(shape, dtype=torch.float32, device="cuda")
            d["um"].add(dw, dst, shape, c_splits[0][1])
            d["dst_ids"].add(id(p))
            self.pgrads[id(p)] = dst
            return
        self._padd(p, ops.unpack_wgrad(dw, shape, c_splits))

    _ARENA_FLOATS = 4 << 20      # 16 MB of fp32 for the many small zero-initialised accumulators of one step
    _ARENA_MAX_ITEM = 1 << 16

    def _zeros(self, *shape, dtype=torch.float32):
        """Zero-initialised buffer.  Small fp32 accumulators (statistics, per-channel sums) are carved out of one
        arena that is cleared with a single memset at the start of a step instead of one fill kernel each."""
        n = 1
        for d in shape:
            n *= d
        if dtype == torch.float32 and n > self._ARENA_MAX_ITEM and self._arena is not None:
            # large fp32 accumulators (the weight-gradient buffers: ~160 per step, 240 MB in total): a second arena, sized by what
            # the first step asked for, cleared by one memset instead of one fill kernel per buffer
            n_al = (n + 31) // 32 * 32
            if self._big is not None and self._big_off + n_al <= self._big.numel():
                t = self._big[self._big_off:self._big_off + n].view(*shape)
                self._big_off += n_al
                self._big_hi = max(self._big_hi, self._big_off)
                return t
            self._big_need += n_al
            return torch.zeros(*shape, dtype=dtype, device="cuda")
        if dtype != torch.float32 or n > self._ARENA_MAX_ITEM or self._arena is None:
            return torch.zeros(*shape, dtype=dtype, device="cuda")
        n_al = (n + 31) // 32 * 32
        if self._arena_off + n_al > self._ARENA_FLOATS:
            return torch.zeros(*shape, dtype=dtype, device="cuda")
        t = self._arena[self._arena_off:self._arena_off + n].view(*shape)
        self._arena_off += n_al
        self._arena_hi = max(self._arena_hi, self._arena_off)
        return t

    # ------------------------------------------------------------------ packed parameters (no BatchNorm folding)
    def _pack_conv_train(self, conv, c_splits, n_pad=None):
        key = (id(conv), "train", str(c_splits), n_pad)

        def build():
            w = conv.weight.detach()
            npad = n_pad or ops.round_up(w.shape[0], 16)
            dev_ok = w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.shape[1] * w.shape[2] * w.shape[3] * 4 <= 48 * 1024
            if dev_ok:
                wp = ops.pack_conv_weight_dev(w, c_splits, npad, self.dt)
                self.launches += 1
            else:
                wp = ops.pack_conv_weight(w.float(), c_splits, npad, self.dt).cuda()
            bp = ops.pad_vec(conv.bias.detach().float(), npad).cuda() if conv.bias is not None else None
            self._pack_recs.pop(key, None)
            if dev_ok and len(c_splits) <= 2 and (bp is None or conv.bias.is_cuda):
                # remembered for repack_all(): every operand of the step re-packed by ONE launch (+ one multi-tensor copy for
                # the padded bias vectors) instead of one launch per conv
                lo1 = c_splits[0][0]
                f = [sum(p for _, p in c_splits), len(c_splits), 0, c_splits[0][0], c_splits[0][1],
                     lo1 if len(c_splits) > 1 else 0, c_splits[1][0] if len(c_splits) > 1 else 0]
                self._pack_recs[key] = dict(kind=0, w=conv.weight, out=wp, rows=npad, cout=w.shape[0], cin=w.shape[1],
                                            kk=w.shape[2] * w.shape[3], f=f, params=[conv.weight, conv.bias], val=(wp, bp),
                                            vec=[(conv.bias, bp)] if bp is not None else [])
            return wp, bp

        return self._pack(key, [conv.weight, conv.bias], build)

    def repack_all(self, invalidate_others=True):
        """Refresh every remembered packed operand from the current parameter values with one prn_pack_multi launch and one
        multi-tensor copy, writing into the SAME buffers (captured graphs and cached descriptors stay valid), and mark the
        cache entries current.  Entries that are not covered (a handful of small per-module packs) are dropped so that their
        builders run again.  Returns False when nothing has been remembered yet (first step)."""
        recs = self._pack_recs
        if not recs:
            return False
        import struct
        sig = tuple((r["w"].data_ptr(), r["out"].data_ptr()) + tuple(q.data_ptr() for pb in r["vec"] for q in pb) for r in recs.values())
        if self._pack_tab is None or self._pack_tab[0] != sig:
            assert not torch.cuda.is_current_stream_capturing(), "call repack_all() once outside the capture (it uploads its tables)"
            blob, work, smem = b"", [], 0
            for i, r in enumerate(recs.values()):
                blob += struct.pack("<QQ12i", r["w"].data_ptr(), r["out"].data_ptr(), r["kind"], r["cout"], r["cin"], r["kk"], *r["f"], r["rows"])
                work += [(i, n) for n in range(0, r["rows"], 8)]          # a block takes 8 consecutive rows (kPackRows)
                if r["kind"] == 0:
                    smem = max(smem, r["cin"] * r["kk"] * 4)
            rec_t = torch.frombuffer(bytearray(blob), dtype=torch.uint8).cuda()
            work_t = torch.tensor(work, dtype=torch.int32).cuda().contiguous()
            # the padded bias vectors: one prn_copy_multi_f32 launch (torch._foreach_copy_ issued one cudaMemcpy per vector: ~200)
            dst, src = [], []
            for r in recs.values():
                for param, buf in r["vec"]:
                    dst.append(buf[:param.numel()])
                    src.append(param.detach())
            vec = None
            if dst:
                vec = ops.CopyMulti(dst)
                vec.set_sources(src)
            self._pack_tab = (sig, rec_t, work_t, len(work), smem, vec)
        _, rec_t, work_t, nblk, smem, vec = self._pack_tab
        L.check(self.lib.prn_pack_multi(C.c_void_p(rec_t.data_ptr()), C.c_void_p(work_t.data_ptr()), nblk, smem, self.dt, self._st()),
                "prn_pack_multi")
        self.launches += 1
        if vec is not None:
            vec.run()
            self.launches += 1
        if invalidate_others:
            for k in [k for k in self._packed if k not in recs]:
                del self._packed[k]
        for k, r in recs.items():
            self._packed[k] = (_ver(*r["params"]), r["val"])
        return True

    def _pack_dgrad(self, conv, real_lo, real_hi, rows_pad, cout_pad):
        """Weights of the input-gradient contraction for input channels [real_lo, real_hi) padded to rows_pad rows;
        K = (ky, kx, cout padded to cout_pad), taps flipped."""
        key = (id(conv), "dgrad", real_lo, real_hi, rows_pad, cout_pad)

        def build():
            w = conv.weight.detach()
            self._pack_recs.pop(key, None)
            if w.is_cuda and w.dtype == torch.float32 and w.is_contiguous():
                self.launches += 1
                out = ops.pack_dgrad_weight_dev(w, real_lo, real_hi, rows_pad, cout_pad, self.dt)
                self._pack_recs[key] = dict(kind=1, w=conv.weight, out=out, rows=rows_pad, cout=w.shape[0], cin=w.shape[1],
                                            kk=w.shape[2] * w.shape[3], f=[real_lo, real_hi - real_lo, cout_pad, rows_pad, 0, 0, 0],
                                            params=[conv.weight], val=out, vec=[])
                return out
            return ops.pack_dgrad_weight(w.float()[:, real_lo:real_hi], cout_pad=cout_pad, n_pad=rows_pad, dtype=self.dt).cuda()

        return self._pack(key, [conv.weight], build)

    # ------------------------------------------------------------------ convolution with tape
    def t_conv(self, x, conv, *, src1=None, c0=None, c_splits=None, stride=None, pad=None, pad_mode=L.PAD_ZERO, upsample=1,
               act=L.ACT_NONE, residual=None, stats=None, stats_cg=0, out16_buf=None, out32_buf=None, out_img_rows=0,
               need_dx=True, token=None):
        """conv (+bias) (+residual) (+ReLU); returns the 16-bit output (or `token` when the output lives in a caller
        buffer).  Gradient of the output is looked up under the returned object."""
        assert act in (L.ACT_NONE, L.ACT_RELU)
        B, H, W, _ = x.shape
        k = conv.kernel_size[0]
        stride = conv.stride[0] if stride is None else stride
        pad = conv.padding[0] if pad is None else pad
        cin0 = c0 if c0 is not None else x.shape[-1]
        real = conv.in_channels
        if c_splits is None:
            c_splits = [(real, cin0)] if src1 is None else [(cin0, cin0), (real - cin0, src1.shape[-1])]
        wp, bp = self._pack_conv_train(conv, c_splits)
        n_pad = wp.shape[0]
        cout = conv.out_channels
        Ho = (H * upsample + 2 * pad - k) // stride + 1
        Wo = (W * upsample + 2 * pad - k) // stride + 1
        o16 = out16_buf if out16_buf is not None else (self._empty(B, Ho, Wo, n_pad) if out32_buf is None or token is None else None)
        flops = 2.0 * B * Ho * Wo * cout * real * k * k
        with self._timed(f"conv{k}x{k}", flops):
            ops.conv2d(x, wp, batch=B, h_in=H, w_in=W, ksize=k, stride=stride, pad=pad, pad_mode=pad_mode, upsample=upsample,
                       src1=src1, bias=bp, residual=residual, act=act, out16=o16, out32=out32_buf, out_img_rows=out_img_rows,
                       stats=stats, stats_cg=stats_cg, dtype=self.dt, c0=c0,
                       ld_out16=(out16_buf.shape[-1] if out16_buf is not None else None),
                       ld_out32=(out32_buf.shape[-1] if out32_buf is not None else None))
        out = token if token is not None else o16
        if token is not None:
            self._keep.append(token)

        def bwd():
            dy = self._take(out)
            if dy is None:
                return
            if act == L.ACT_RELU:
                g = torch.empty_like(dy)
                ops.relu_bwd(dy, o16, g, self.dt)
                self.launches += 1
                dy = g
            if residual is not None:
                self._give(residual, dy, owned=False)
            ldy = dy.shape[-1]
            if conv.bias is not None and conv.bias.requires_grad:
                sums = self._zeros(ldy, 2)
                ops.chan_reduce(dy, None, None, None, sums, self.dt)
                self.launches += 1
                self._padd(conv.bias, sums[:cout, 0])
            if conv.weight.requires_grad:
                dw = self._wbuf(conv, c_splits)
                self._wgrad(f"wgrad{k}x{k}", flops, x, dy, dw, batch=B, h_in=H, w_in=W, n=cout, ksize=k, stride=stride, pad=pad,
                            pad_mode=pad_mode, upsample=upsample, src1=src1, c0=c0, dtype=self.dt)
            if need_dx:
                assert ldy % 64 == 0, "input-gradient contraction needs dY rows padded to a multiple of 64 channels"
                lo = 0
                for src, (real_c, pad_c) in zip([x, src1], c_splits):
                    self._dgrad(src, conv, dy, lo, lo + real_c, pad_c, k, stride, pad, pad_mode, upsample, (B, H, W), flops * real_c / real)
                    lo += real_c

        self.tape.append(bwd)
        return out

    def _wgrad(self, name, flops, src0, dy, dw, **kw):
        """Weight-gradient contraction.  It depends on (activation, dY) only and nothing in the backward chain depends
        on it, so it goes to a side stream where its CTAs fill the SMs that the (often sub-wave) input-gradient and
        normalisation kernels of the main stream leave idle; joined before the accumulators are unpacked."""
        if not self.side_wgrad or self.profile is not None:
            with self._timed(name, flops):
                ops.conv2d_wgrad(src0, dy, dw, **kw)
            return
        main = torch.cuda.current_stream()
        if self._wg_streams is None:
            self._wg_streams = [torch.cuda.Stream() for _ in range(2)]
        st = self._wg_streams[self._wg_rr % len(self._wg_streams)]
        self._wg_rr += 1
        ev = torch.cuda.Event()
        ev.record(main)
        st.wait_event(ev)
        with torch.cuda.stream(st):
            ops.conv2d_wgrad(src0, dy, dw, **kw)
        for t in (src0, dy, dw, kw.get("src1")):
            if t is not None:
                t.record_stream(st)
        self._wg_used.add(st)
        self.launches += 1

    def _join_wgrad(self):
        main = torch.cuda.current_stream()
        for st in self._wg_used:
            ev = torch.cuda.Event()
            ev.record(st)
            main.wait_event(ev)
        self._wg_used = set()

    def _wbuf(self, conv, c_splits):
        key = (id(conv), str(c_splits))
        dw = self.wbufs.get(key)
        if dw is None:
            k = conv.kernel_size[0]
            cpad = sum(p for _, p in c_splits)
            dw = self._zeros(ops.round_up(conv.out_channels, 4), k * k * cpad)
            self.wbufs[key] = dw

            def fin():
                self._fin_wgrad(conv.weight, dw, tuple(conv.weight.shape), c_splits)

            self.pfinal.append(fin)
        return dw

    def _dgrad(self, src, conv, dy, lo, hi, pad_c, k, stride, pad, pad_mode, upsample, in_shape, flops):
        """Gradient w.r.t. channels [lo, hi) of the conv input living in the first pad_c channels of `src`."""
        B, H, W = in_shape
        width = src.shape[-1]
        wd = self._pack_dgrad(conv, lo, hi, pad_c, dy.shape[-1])
        He, We = H * upsample, W * upsample
        cur = self._peek(src)
        if pad_mode == L.PAD_REFLECT:
            assert k == 3 and pad == 1 and stride == 1 and pad_c == width
            dpad = self._empty(B, He + 2, We + 2, pad_c)
            with self._timed("dgrad3x3", flops):
                ops.conv2d(dy, wd, batch=B, h_in=He, w_in=We, ksize=3, stride=1, pad=2, out16=dpad, dtype=self.dt)
            if cur is None:
                din = self._empty(B, H, W, width)
                ops.reflect_fold(dpad, din, upsample, False, self.dt)
                self._set_grad(src, din)
            else:
                din = self._own(src)
                ops.reflect_fold(dpad, din, upsample, True, self.dt)
            self.launches += 1
            return
        assert upsample == 1
        if stride == 1:
            if cur is None:
                din = self._zeros(B, H, W, width, dtype=self.tdt) if pad_c < width else self._empty(B, H, W, width)
                res = None
            else:
                din = self._own(src)
                res = din
            with self._timed(f"dgrad{k}x{k}", flops):
                ops.conv2d(dy, wd, batch=B, h_in=H, w_in=W, ksize=k, stride=1, pad=k - 1 - pad, residual=res,
                           ld_res=width if res is not None else None, out16=din, ld_out16=width, dtype=self.dt)
            if cur is None:
                self._set_grad(src, din)
            return
        assert k == 1 and pad == 0 and pad_c == width, "strided input gradients exist for the 1x1 projection only"
        Ho, Wo = dy.shape[1], dy.shape[2]
        tmp = self._empty(B, Ho, Wo, width)
        with self._timed("dgrad1x1", flops):
            ops.conv2d(dy, wd, batch=B, h_in=Ho, w_in=Wo, ksize=1, out16=tmp, dtype=self.dt)
        if cur is None:
            din = self._zeros(B, H, W, width, dtype=self.tdt)
            self._set_grad(src, din)
        else:
            din = self._own(src)
        ops.add_strided(din, tmp, stride, self.dt)
        self.launches += 1

    def _own(self, t):
        """The gradient buffer of t, made writable (copied once if it is shared with another activation)."""
        e = self.grads[id(t)]
        if not e[1]:
            e[0] = e[0].clone()
            e[1] = True
        return e[0]

    # ------------------------------------------------------------------ BatchNorm (batch statistics) [+ residual] [+ ReLU]
    def t_bn(self, y, bn, relu, residual=None):
        """y: conv output whose epilogue accumulated per-channel {sum, sumsq} into y._prn_stats."""
        Cc = y.shape[-1]
        rows = y.numel() // Cc
        gamma, beta = bn.weight, bn.bias
        gb = self._pack((id(bn), "bn_gb"), [gamma, beta], lambda: (ops.pad_vec(gamma, Cc).cuda(), ops.pad_vec(beta, Cc).cuda()))
        if bn.training:
            mi = self._empty(Cc, 2, dtype=torch.float32)
            rm = bn.running_mean if bn.track_running_stats else None
            rv = bn.running_var if bn.track_running_stats else None
            assert rm is None or rm.numel() == Cc
            mom = bn.momentum if bn.momentum is not None else 0.1
            fused_finalize = rm is None or rm.numel() == Cc
            if not fused_finalize:
                ops.bn_finalize(y._prn_stats, mi, rm, rv, rows, bn.eps, mom)
                self.launches += 1
            if bn.track_running_stats:
                self._nbt.append(bn.num_batches_tracked)      # advanced with one multi-tensor add at the end of the forward
                # the kernel wrote through raw pointers: bump the version counters the weight caches key on
                torch.autograd.graph.increment_version(bn.running_mean)
                torch.autograd.graph.increment_version(bn.running_var)
        else:   # frozen statistics (freeze_bn): normalise with the running estimates
            mi = self._pack((id(bn), "bn_frozen"), [bn.running_mean, bn.running_var],
                            lambda: torch.stack([ops.pad_vec(bn.running_mean, Cc),
                                                 1.0 / torch.sqrt(ops.pad_vec(bn.running_var, Cc) + bn.eps)], 1).contiguous().cuda())
        out = self._empty(*y.shape)
        if bn.training and fused_finalize:      # statistics finalised inside the apply pass (one launch less per BatchNorm)
            ops.bn_finalize_apply(y, out, y._prn_stats, mi, rm, rv, rows, bn.eps, mom, gb[0], gb[1], residual, relu, self.dt)
        else:
            ops.bn_apply(y, out, mi, gb[0], gb[1], residual, relu, self.dt)
        self.launches += 1
        train_stats = bn.training
        nreal = gamma.numel()

        def bwd():
            dz = self._take(out)
            if dz is None:
                return
            sums = self._zeros(Cc, 2)
            need_param = gamma.requires_grad or beta.requires_grad
            if train_stats or need_param:
                ops.chan_reduce(dz, out if relu else None, y, mi, sums, self.dt)
                self.launches += 1
            if need_param:
                self._padd(gamma, sums[:nreal, 1])
                self._padd(beta, sums[:nreal, 0])
            if not train_stats:
                sums = self._zeros(Cc, 2)      # frozen statistics: dx = gamma * invstd * g
            dx = self._empty(*y.shape)
            gout = self._empty(*y.shape) if residual is not None else None
            ops.bn_bwd_apply(dz, out if relu else None, y, mi, gb[0], sums, dx, gout, self.dt)
            self.launches += 1
            self._give(y, dx)
            if residual is not None:
                self._give(residual, gout)

        self.tape.append(bwd)
        return out

    def conv_bn(self, x, conv, bn, relu, residual=None, **kw):
        n_pad = ops.round_up(conv.out_channels, 16)
        stats = self._zeros(n_pad, 2)
        y = self.t_conv(x, conv, stats=stats, stats_cg=0, **kw)
        y._prn_stats = stats
        return self.t_bn(y, bn, relu, residual)

    # ------------------------------------------------------------------ conv -> GroupNorm(32) -> ReLU
    def conv_gn_relu_t(self, x, conv, gn, **kw):
        B = x.shape[0]
        G = gn.num_groups
        cg = conv.out_channels // G
        stats = self._zeros(B * G * 2)
        y = self.t_conv(x, conv, stats=stats, stats_cg=cg, **kw)
        out = self.gn_relu(y, stats, gn)
        gamma, beta = gn.weight, gn.bias
        gd = self._pack((id(gn), "gn"), [gamma, beta], lambda: (gamma.detach().float().cuda().contiguous(), beta.detach().float().cuda().contiguous()))[0]
        key = (id(gn), "dgb")
        Cc = y.shape[-1]

        def bwd():
            dz = self._take(out)
            if dz is None:
                return
            dgb = self.wbufs.get(key)
            if dgb is None:
                dgb = self.wbufs[key] = self._zeros(Cc, 2)

                def fin():
                    self._padd(gamma, dgb[:, 1])
                    self._padd(beta, dgb[:, 0])

                self.pfinal.append(fin)
            sums_bc = self._zeros(B, Cc, 2)
            ops.gn_bwd_reduce(dz, out, y, stats, sums_bc, dgb, cg, gn.eps, self.dt)
            dx = self._empty(*y.shape)
            ops.gn_bwd_apply(dz, out, y, stats, gd, sums_bc, dx, cg, gn.eps, self.dt)
            self.launches += 2
            self._give(y, dx)

        self.tape.append(bwd)
        return out

    # ------------------------------------------------------------------ modulated deformable 3x3 (training form)
    def dcn_t(self, x, m):
        """models/dcn.py:52-67 -> pre-BatchNorm output with per-channel sums attached."""
        B, H, W, Cc = x.shape
        stride = m.stride if isinstance(m.stride, int) else m.stride[0]
        pad = m.padding if isinstance(m.padding, int) else m.padding[0]
        reg = m.regular_conv
        N = reg.out_channels
        bound = max(H, W) / 4.0

        def build_om():
            w = torch.cat([m.offset_conv.weight.detach().float(), m.modulator_conv.weight.detach().float()], 0)
            b = torch.cat([m.offset_conv.bias.detach().float(), m.modulator_conv.bias.detach().float()], 0)
            wd = ops.pack_dgrad_weight(w, cout_pad=64, n_pad=Cc, dtype=self.dt).cuda()
            return ops.pack_conv_weight(w, [(w.shape[1], Cc)], 32, self.dt).cuda(), ops.pad_vec(b, 32).cuda(), wd

        om_params = [m.offset_conv.weight, m.offset_conv.bias, m.modulator_conv.weight, m.modulator_conv.bias]
        wom, bom, wom_d = self._pack((id(m), "offmask_train"), om_params, build_om)
        Ho = (H + 2 * pad - 3) // stride + 1
        Wo = (W + 2 * pad - 3) // stride + 1
        om = self._empty(B, Ho, Wo, 32, dtype=torch.float32)
        with self._timed("conv3x3", 2.0 * B * Ho * Wo * 27 * Cc * 9):
            ops.conv2d(x, wom, batch=B, h_in=H, w_in=W, ksize=3, stride=stride, pad=pad, bias=bom, act=L.ACT_DCN_OFFMASK,
                       act_param=bound, out32=om, dtype=self.dt)
        col = self._empty(B, Ho, Wo, 9 * Cc)
        ops.dcn_im2col(x, om, col, stride, pad, self.dt)
        self.launches += 1

        def build_reg():
            w = reg.weight.detach().float()
            wp = ops.pack_conv_weight(w, [(w.shape[1], Cc)], ops.round_up(N, 16), self.dt).cuda()      # [N, 9*C], k = (tap, c)
            bp = ops.pad_vec(reg.bias.detach().float(), wp.shape[0]).cuda() if reg.bias is not None else None
            return wp, bp, wp[:N].t().contiguous()                                                      # [9*C, N]

        wreg, breg, wreg_t = self._pack((id(reg), "dcn_reg_train"), [reg.weight, reg.bias], build_reg)
        n_pad = wreg.shape[0]
        stats = self._zeros(n_pad, 2)
        y = self._empty(B, Ho, Wo, n_pad)
        flops = 2.0 * B * Ho * Wo * N * Cc * 9
        with self._timed("dcn", flops):
            ops.conv2d(col, wreg, batch=B, h_in=Ho, w_in=Wo, ksize=1, bias=breg, out16=y, stats=stats, stats_cg=0, dtype=self.dt)
        y._prn_stats = stats

        def bwd():
            dy = self._take(y)
            if dy is None:
                return
            if reg.bias is not None and reg.bias.requires_grad:
                sums = self._zeros(n_pad, 2)
                ops.chan_reduce(dy, None, None, None, sums, self.dt)
                self._padd(reg.bias, sums[:N, 0])
            dw = self._zeros(ops.round_up(N, 4), 9 * Cc)
            self._wgrad("wgrad_dcn", flops, col, dy, dw, batch=B, h_in=Ho, w_in=Wo, n=N, ksize=1, dtype=self.dt)
            self.pfinal.append(lambda: self._fin_wgrad(reg.weight, dw, tuple(reg.weight.shape), [(reg.in_channels, Cc)]))
            dcol = self._empty(B, Ho, Wo, 9 * Cc)
            with self._timed("dgrad_dcn", flops):
                ops.conv2d(dy, wreg_t, batch=B, h_in=Ho, w_in=Wo, ksize=1, c0=N, out16=dcol, dtype=self.dt)
            dx32 = self._zeros(B, H, W, Cc)
            dpre = self._empty(B, Ho, Wo, 64)
            ops.dcn_col2im_bwd(x, om, dcol, dx32, dpre, stride, pad, bound, self.dt)
            # offset / modulator conv: bias, weight and input gradients
            sums = self._zeros(64, 2)
            ops.chan_reduce(dpre, None, None, None, sums, self.dt)
            self._padd(m.offset_conv.bias, sums[:18, 0])
            self._padd(m.modulator_conv.bias, sums[18:27, 0])
            dwom = self._zeros(28, 9 * Cc)
            self._wgrad("wgrad3x3", 2.0 * B * Ho * Wo * 27 * Cc * 9, x, dpre, dwom, batch=B, h_in=H, w_in=W, n=27, ksize=3,
                        stride=stride, pad=pad, dtype=self.dt)

            def fin_om():
                g27 = ops.unpack_wgrad(dwom, (27, m.offset_conv.in_channels, 3, 3), [(m.offset_conv.in_channels, Cc)])
                self._padd(m.offset_conv.weight, g27[:18])
                self._padd(m.modulator_conv.weight, g27[18:27])

            self.pfinal.append(fin_om)
            if stride == 1:
                u = dpre
            else:   # transposed stride-2 conv = stride-1 conv over the zero-inserted gradient
                u = self._zeros(B, H, W, 64, dtype=self.tdt)
                ops.add_strided(u, dpre, stride, self.dt)
            t16 = self._empty(B, H, W, Cc)
            ops.conv2d(u, wom_d, batch=B, h_in=H, w_in=W, ksize=3, stride=1, pad=3 - 1 - pad, out16=t16, dtype=self.dt)
            da = self._empty(B, H, W, Cc)
            ops.add_f32(dx32, t16, da, self.dt)
            self.launches += 7
            self._give(x, da)

        self.tape.append(bwd)
        return y

    # ------------------------------------------------------------------ pooling / resampling with tape
    def maxpool_t(self, x):
        out = self.maxpool(x)

        def bwd():
            dz = self._take(out)
            if dz is None:
                return
            din = self._empty(*x.shape)
            ops.maxpool_bwd(x, dz, din, self.dt)
            self.launches += 1
            self._give(x, din)

        self.tape.append(bwd)
        return out

    def avgpool2_t(self, x):
        out = self.avgpool2(x)

        def bwd():
            dz = self._take(out)
            if dz is None:
                return
            if self._peek(x) is None:
                din = self._empty(*x.shape)
                ops.avgpool2_bwd(dz, din, False, self.dt)
                self._set_grad(x, din)
            else:
                ops.avgpool2_bwd(dz, self._own(x), True, self.dt)
            self.launches += 1

        self.tape.append(bwd)
        return out

    def upsample2x_t(self, x, into=None):
        """Bilinear x2; `into` accumulates in place (its gradient is looked up under the `into` tensor)."""
        out = self.upsample2x(x, into=into)

        def bwd():
            dz = self._peek(out) if into is not None else self._take(out)
            if dz is None:
                return
            din = self._empty(*x.shape)
            ops.upsample2x_bwd(dz, din, self.dt)
            self.launches += 1
            self._give(x, din)

        self.tape.append(bwd)
        return out

    def resize_coord_t(self, f, S, c_out):
        B, h, w, cf = f.shape
        kf = self._empty(B, S, S, c_out)
        self._call(self.lib.prn_resize_bilinear, C.c_void_p(f.data_ptr()), C.c_void_p(kf.data_ptr()), B, h, w, cf, S, S, c_out, 1,
                   self.dt, self._st())

        def bwd():
            dz = self._take(kf)
            if dz is None:
                return
            d32 = self._zeros(B, h, w, cf)
            ops.resize_bilinear_bwd(dz, d32, self.dt)
            din = self._empty(B, h, w, cf)
            ops.add_f32(d32, None, din, self.dt)
            self.launches += 2
            self._give(f, din)

        self.tape.append(bwd)
        return kf

    def append_coord_t(self, x, c_out):
        B, h, w, cf = x.shape
        xc = self._empty(B, h, w, c_out)
        self._call(self.lib.prn_append_coord, C.c_void_p(x.data_ptr()), C.c_void_p(xc.data_ptr()), B, h, w, cf, c_out, self.dt, self._st())

        def bwd():
            dz = self._take(xc)
            if dz is None:
                return
            self._give(x, dz[..., :cf].contiguous())

        self.tape.append(bwd)
        return xc

    # ------------------------------------------------------------------ stages
    def bottleneck_t(self, x, blk):
        """models/backbone.py:53-73 with batch-statistics BatchNorm."""
        a1 = self.conv_bn(x, blk.conv1, blk.bn1, True)
        if isinstance(blk.conv2, DeformableConv2d):
            y2 = self.dcn_t(a1, blk.conv2)
            a2 = self.t_bn(y2, blk.bn2, True)
        else:
            a2 = self.conv_bn(a1, blk.conv2, blk.bn2, True)
        res = x
        if blk.downsample is not None:
            res = self.conv_bn(x, blk.downsample[0], blk.downsample[1], False)
        return self.conv_bn(a2, blk.conv3, blk.bn3, True, residual=res)

    def backbone_t(self, x, bb):
        """models/backbone.py:197-209 (training).  x: NCHW fp32 CUDA; no gradient flows to the image."""
        assert x.is_cuda and x.dim() == 4 and x.shape[1] == 3
        x = x.detach().float().contiguous()
        B, _, H, W = x.shape
        a = self._empty(B, H // 2, W // 2, 192)
        self._call(self.lib.prn_stem_im2col, C.c_void_p(x.data_ptr()), C.c_void_p(a.data_ptr()), B, H, W, self.dt, self._st())
        conv1 = bb.conv1

        def build_stem():
            wk = conv1.weight.detach().float().permute(0, 2, 3, 1).reshape(64, 147)
            return torch.nn.functional.pad(wk, (0, 192 - 147)).to(self.tdt).contiguous().cuda()

        wk = self._pack((id(conv1), "stem_train"), [conv1.weight], build_stem)
        stats = self._zeros(64, 2)
        y = self._empty(B, H // 2, W // 2, 64)
        flops = 2.0 * B * (H // 2) * (W // 2) * 64 * 147
        with self._timed("conv7x7", flops):
            ops.conv2d(a, wk, batch=B, h_in=H // 2, w_in=W // 2, ksize=1, out16=y, stats=stats, stats_cg=0, dtype=self.dt)
        y._prn_stats = stats

        def bwd():
            dy = self._take(y)
            if dy is None or not conv1.weight.requires_grad:
                return
            dw = self._zeros(64, 192)
            self._wgrad("wgrad7x7", flops, a, dy, dw, batch=B, h_in=H // 2, w_in=W // 2, n=64, ksize=1, dtype=self.dt)
            self.pfinal.append(lambda: self._padd(conv1.weight, dw[:, :147].reshape(64, 7, 7, 3).permute(0, 3, 1, 2).contiguous()))

        self.tape.append(bwd)
        t = self.maxpool_t(self.t_bn(y, bb.bn1, True))
        outs = []
        for li, layer in enumerate(bb.layers):
            if li == 2:
                self._marks.append(len(self.tape))       # gradient bucket boundary: stem + layers 0-1 | layers 2-3
            for blk in layer:
                t = self.bottleneck_t(t, blk)
            outs.append(t)
        self._marks.append(len(self.tape))               # layers 2-3 | FPN + heads + decoder
        return outs

    def fpn_t(self, cs, fpn):
        """models/fpn.py:45-63."""
        lats, prev = [], None
        for i, c in enumerate(cs):
            res = self.avgpool2_t(prev) if prev is not None else None
            prev = self.t_conv(c, fpn.lateral_convs[i], residual=res)
            lats.append(prev)
        return [self.t_conv(l, fpn.fpn_convs[i], act=L.ACT_RELU) for i, l in enumerate(lats)]

    def inst_head_t(self, feats, head):
        """planerecnet.py:355-391.  Returns the joint buffers plus per-level gradient tokens."""
        B = feats[0].shape[0]
        grids = head.num_grids
        total = sum(s * s for s in grids)
        nk = head.num_kernels
        kern16 = self._empty(B, total, nk)
        kern32 = self._empty(B, total, nk, dtype=torch.float32)
        cate32 = self._empty(B, total, 16, dtype=torch.float32)
        cin = head.instance_in_channels
        offs = [sum(g * g for g in grids[:l]) for l in range(len(grids))]
        ktok, ctok = [], []
        for lvl, f in enumerate(feats):
            S, off = grids[lvl], offs[lvl]
            kf = self.resize_coord_t(f, S, ops.round_up(cin + 2, 64))
            t = kf
            for i in range(0, len(head.kernel_tower), 3):
                t = self.conv_gn_relu_t(t, head.kernel_tower[i], head.kernel_tower[i + 1])
            tok = object()
            self.t_conv(t, head.kernel_pred, out16_buf=kern16[:, off:], out32_buf=kern32[:, off:], out_img_rows=total, token=tok)
            ktok.append(tok)
            t = kf
            for i in range(0, len(head.cate_tower), 3):
                t = self.conv_gn_relu_t(t, head.cate_tower[i], head.cate_tower[i + 1], c0=cin if i == 0 else None)
            tok = object()
            self.t_conv(t, head.cate_pred, out32_buf=cate32[:, off:], out_img_rows=total, token=tok)
            ctok.append(tok)
        return {"kern16": kern16, "kern32": kern32, "cate32": cate32, "ktok": ktok, "ctok": ctok}

    def mask_head_t(self, ps, head):
        """planerecnet.py:467-496 (the level sum is built out of place: level 0's output is its own ReLU mask)."""
        lv = head.convs_all_levels
        lvl0 = self.conv_gn_relu_t(ps[0], lv[0].conv0[0], lv[0].conv0[1])
        acc = None
        for i in range(1, head.num_levels):
            x = ps[i]
            if i == 3:
                x = self.append_coord_t(x, ops.round_up(x.shape[-1] + 2, 64))
            for j in range(i):
                tower = getattr(lv[i], f"conv{j}")
                x = self.conv_gn_relu_t(x, tower[0], tower[1])
                if j < i - 1:
                    x = self.upsample2x_t(x)
                elif acc is None:
                    up = self.upsample2x_t(x)
                    acc = self._empty(*lvl0.shape)
                    ops.add16(lvl0, up, acc, self.dt)
                    self.launches += 1
                    acc_t, lvl0_t, up_t = acc, lvl0, up

                    def bwd_add():      # earliest reader of acc's gradient on the tape = last to run: consumes it
                        dz = self._take(acc_t)
                        if dz is not None:
                            self._give(lvl0_t, dz, owned=False)
                            self._give(up_t, dz, owned=False)

                    self.tape.append(bwd_add)
                else:
                    self.upsample2x_t(x, into=acc)
        return self.conv_gn_relu_t(acc, head.conv_pred[0], head.conv_pred[1])

    def _rconv_t(self, x, seq, src1=None):
        up = 2 if isinstance(seq[0], nn.Upsample) else 1
        conv = next(m for m in seq if isinstance(m, nn.Conv2d))
        bn = next(m for m in seq if isinstance(m, nn.BatchNorm2d))
        return self.conv_bn(x, conv, bn, True, src1=src1, pad=1, pad_mode=L.PAD_REFLECT, upsample=up)

    def depth_decoder_t(self, cs, mask16, kern16, dec):
        """planerecnet.py:586-607.  The plane-prior operands are detached in the reference (:589,:592): only
        conv1x1's weight and bias receive a gradient from that branch."""
        B, mh, mw, mc = mask16.shape
        total = kern16.shape[1]
        kpad = ops.round_up(total, 64)
        q = self._empty(B, (mh // 4) * (mw // 4) * 4, mc)
        self._call(self.lib.prn_ppa_gather, C.c_void_p(mask16.data_ptr()), C.c_void_p(q.data_ptr()), B, mh, mw, mc, self.dt, self._st())
        p = torch.zeros(B, mh // 4, mw // 4, kpad, dtype=self.tdt, device="cuda")
        with self._timed("ppa_dyn", 2.0 * B * (mh // 4) * (mw // 4) * 4 * total * mc):
            ops.conv2d(q, kern16.reshape(B * total, mc), batch=B, h_in=(mh // 4) * (mw // 4), w_in=4, ksize=1,
                       act=L.ACT_SIGMOID_AVG4, out16=p, ld_out16=kpad, n_pad=total, w_group_rows=total,
                       out_img_rows=(mh // 4) * (mw // 4), dtype=self.dt)
        attn = self.t_conv(p, dec.conv1x1[0], c_splits=[(total, kpad)], need_dx=False)

        feats = list(reversed(cs))
        x = self.t_conv(feats[0], dec.latlayer1)
        x = self._rconv_t(x, dec.conv1)
        x = self._rconv_t(x, dec.deconv1)
        skips = {}
        for k in (2, 3, 4):
            sk = self.t_conv(feats[k - 1], getattr(dec, f"latlayer{k}"))
            skips[k] = self._rconv_t(sk, getattr(dec, f"conv{k}"))
        xa = self._empty(*x.shape)
        self._call(self.lib.prn_mul, C.c_void_p(x.data_ptr()), C.c_void_p(attn.data_ptr()), C.c_void_p(xa.data_ptr()),
                   C.c_int64(x.numel()), self.dt, self._st())
        x0 = x

        def bwd_mul():
            dz = self._take(xa)
            if dz is None:
                return
            d_x = self._empty(*x0.shape)
            d_a = self._empty(*x0.shape)
            for a_, b_, o_ in ((dz, attn, d_x), (dz, x0, d_a)):
                self._call(self.lib.prn_mul, C.c_void_p(a_.data_ptr()), C.c_void_p(b_.data_ptr()), C.c_void_p(o_.data_ptr()),
                           C.c_int64(o_.numel()), self.dt, self._st())
            self._give(x0, d_x)
            self._give(attn, d_a)

        self.tape.append(bwd_mul)
        x = self._rconv_t(x0, dec.refine_conv, src1=xa)
        for k in (2, 3, 4):
            x = self._rconv_t(skips[k], getattr(dec, f"deconv{k}"), src1=x)
        # depth head 64 -> 1 (+ softplus) on the CUDA cores; its backward goes through the generic contractions
        dconv = dec.depth_pred[1]
        w9c = self._pack((id(dconv), "to1_train"), [dconv.weight, dconv.bias],
                         lambda: (dconv.weight.detach().float()[0].permute(1, 2, 0).reshape(9, -1).contiguous().cuda(),
                                  dconv.bias.detach().float().reshape(1).contiguous().cuda()))
        Bx, Hx, Wx, Cx = x.shape
        d32 = self._empty(Bx, Hx, Wx, 1, dtype=torch.float32)
        self._call(self.lib.prn_conv3x3_to1_reflect_devbias, C.c_void_p(x.data_ptr()), C.c_void_p(w9c[0].data_ptr()),
                   C.c_void_p(w9c[1].data_ptr()), C.c_void_p(d32.data_ptr()), Bx, Hx, Wx, Cx, 1, self.dt, self._st())
        xin = x
        tok = object()
        self._keep.append(tok)

        def bwd_head():
            dout = self._take(tok)      # fp32 [B,1,H,W] == [B,H,W,1]
            if dout is None:
                return
            dpre = self._empty(Bx, Hx, Wx, 64)
            ops.softplus_bwd_pad(dout, d32, dpre, self.dt)
            sums = self._zeros(64, 2)
            ops.chan_reduce(dpre, None, None, None, sums, self.dt)
            self._padd(dconv.bias, sums[:1, 0])
            dw = self._zeros(4, 9 * Cx)
            self._wgrad("wgrad3x3", 2.0 * Bx * Hx * Wx * Cx * 9, xin, dpre, dw, batch=Bx, h_in=Hx, w_in=Wx, n=1, ksize=3, stride=1,
                        pad=1, pad_mode=L.PAD_REFLECT, dtype=self.dt)
            self.pfinal.append(lambda: self._fin_wgrad(dconv.weight, dw, tuple(dconv.weight.shape), [(Cx, Cx)]))
            self.launches += 3
            self._dgrad(xin, dconv, dpre, 0, Cx, Cx, 3, 1, 1, L.PAD_REFLECT, 1, (Bx, Hx, Wx), 2.0 * Bx * Hx * Wx * Cx * 9)

        self.tape.append(bwd_head)
        return d32, tok

    # ------------------------------------------------------------------ whole step
    def forward_train(self, net, x):
        """planerecnet.py:73-103 (training branch).  Returns (mask_pred, [cate]x4, [kernel]x4, depth) NCHW fp32."""
        if not x.is_cuda:
            raise L.PrnError("PlaneRecNet (B200) forward needs a CUDA input tensor; there is no CPU path")
        self.reset()
        cs_all = self.backbone_t(x, net.backbone)
        ps = self.fpn_t([cs_all[i] for i in net.fpn_indices], net.fpn)
        feats = [self.avgpool2_t(ps[0]), ps[1], ps[2], ps[3]]
        inst = self.inst_head_t(feats, net.inst_head)
        mask16 = self.mask_head_t(ps, net.mask_head)
        d32, dtok = self.depth_decoder_t([cs_all[i] for i in net.depth_decoder_indices], mask16, inst["kern16"], net.depth_decoder)
        cates, kerns = self.inst_outputs_nchw(inst, net.inst_head)
        self._flush_nbt()
        self._out = {"mask16": mask16, "ktok": inst["ktok"], "ctok": inst["ctok"], "dtok": dtok}
        return (self.to_nchw(mask16, net.num_masks), cates, kerns, self.to_nchw(d32, 1))

    def _flush_nbt(self):
        if self._nbt:
            torch._foreach_add_(self._nbt, 1)
            self._nbt = []

    def reset(self):
        """Start of a step: drop the previous tape and clear the accumulator arena (one memset)."""
        self._nbt = []
        self.tape, self.grads, self.wbufs, self.pgrads, self.pfinal, self._keep = [], {}, {}, {}, [], []
        self._marks, self._pf_done = [], 0
        if self._arena is None:
            self._arena = torch.zeros(self._ARENA_FLOATS, dtype=torch.float32, device="cuda")
            self._arena_hi = 0
        elif self._arena_hi > 0:
            self._arena[:self._arena_hi].zero_()
        self._arena_off = 0
        if self._big is None:
            # allocated once the first step has shown how much is needed — never inside a graph capture (the buffer must outlive
            # every graph that carves from it)
            if self._big_need > 0 and not torch.cuda.is_current_stream_capturing():
                self._big = torch.zeros(self._big_need, dtype=torch.float32, device="cuda")
                self._big_hi = self._big_need     # sized by one step's needs: every later step clears all of it (a capture that
                                                  # starts right after the allocation must record that memset too)
        elif self._big_hi > 0:
            self._big[:self._big_hi].zero_()
        self._big_off, self._big_need = 0, 0

    def seed_output_grads(self, d_mask, d_cates, d_kerns, d_depth):
        """Cotangents of the training outputs (NCHW fp32 or None) -> NHWC 16-bit gradients of the dense tensors."""
        o = self._out
        if self.grad_scale != 1.0:
            sc = self.grad_scale
            d_mask = None if d_mask is None else d_mask * sc
            d_cates = [None if d is None else d * sc for d in d_cates]
            d_kerns = [None if d is None else d * sc for d in d_kerns]
            d_depth = None if d_depth is None else d_depth * sc
        if d_mask is not None:
            self._set_grad(o["mask16"], self.to_nhwc(d_mask, c_pad=o["mask16"].shape[-1]))
        for tok, d in zip(o["ktok"], d_kerns):
            if d is not None:
                self._set_grad(tok, self.to_nhwc(d, c_pad=ops.round_up(d.shape[1], 64)))
        for tok, d in zip(o["ctok"], d_cates):
            if d is not None:
                self._set_grad(tok, self.to_nhwc(d, c_pad=64))
        if d_depth is not None:
            self._set_grad(o["dtok"], d_depth.detach().float().contiguous())

    def backward_keep_inputs(self, *inputs):
        """Like backward(), for partial graphs (tests, per-module use): also returns the gradient of `inputs`."""
        for fn in reversed(self.tape):
            fn()
        self._join_wgrad()
        for fn in self.pfinal:
            fn()
        dxs = [self._take(t) for t in inputs]
        out = {"params": self.pgrads, "dx": dxs[0] if len(dxs) == 1 else dxs}
        self.tape, self.grads, self.wbufs, self.pfinal, self._keep = [], {}, {}, [], []
        return out

    def backward(self):
        """Replay the tape in reverse; returns {id(param): fp32 gradient}."""
        for fn in reversed(self.tape):
            fn()
        self._join_wgrad()
        for fn in self.pfinal[self._pf_done:]:
            fn()
        if self._defer is not None and self._defer["um"].items:
            self._defer["um"].run()                     # every recorded weight gradient, accumulator -> parameter layout, one launch
            self.launches += 1
        grads = self.pgrads
        self.tape, self.grads, self.wbufs, self.pfinal, self._keep = [], {}, {}, [], []
        self._marks, self._pf_done = [], 0
        return grads

    # Gradient buckets for the data-parallel all-reduce (reverse execution order, decoder / heads first, backbone
    # last): the tape is cut where forward_train entered layers 2-3 and where it left the backbone.  A bucket's parameter
    # gradients are final once its tape segment has been replayed, its side-stream weight gradients joined and its accumulators
    # unpacked — `backward_segment` does exactly that, so the caller can start the bucket's all-reduce while the next segment runs.
    N_BUCKETS = 3

    def bucket_of_params(self, net):
        """{id(param): bucket}: 0 = FPN + heads + depth decoder (first to finish), 1 = backbone layers 2-3, 2 = stem + layers 0-1."""
        out = {}
        for name, p in net.named_parameters():
            if not name.startswith("backbone."):
                out[id(p)] = 0
            elif name.startswith("backbone.layers.2.") or name.startswith("backbone.layers.3."):
                out[id(p)] = 1
            else:
                out[id(p)] = 2
        return out

    def backward_segment(self, b):
        """Replay tape segment b (0, 1, 2 in this order), join the weight-gradient streams and unpack the accumulators that
        exist so far.  After segment b the gradients of bucket b's parameters are final.  Returns self.pgrads (live dict)."""
        assert len(self._marks) == 2, "forward_train did not record the bucket boundaries"
        hi = [len(self.tape), self._marks[1], self._marks[0]][b]
        lo = [self._marks[1], self._marks[0], 0][b]
        for fn in reversed(self.tape[lo:hi]):
            fn()
        self._join_wgrad()
        for fn in self.pfinal[self._pf_done:]:
            fn()
        self._pf_done = len(self.pfinal)
        if b == self.N_BUCKETS - 1:
            grads = self.pgrads
            self.tape, self.grads, self.wbufs, self.pfinal, self._keep = [], {}, {}, [], []
            self._marks, self._pf_done = [], 0
            return grads
        return self.pgrads


@contextlib.contextmanager
def _no_gc():
    """Stream capture runs in global error mode: a cyclic-GC pass in the middle of it may destroy CUDA graphs / events
    of earlier steps (cudaGraphExecDestroy etc.), which invalidates the capture.  Collect first, then keep the GC off."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        yield
    finally:
        if was:
            gc.enable()


def _allreduce_mean(t, world):
    """NCCL mean over the ranks.  PRN_DP_OP=avg uses ReduceOp.AVG (one kernel); the default is SUM followed by an in-place scale."""
    import os
    import torch.distributed as dist
    if os.environ.get("PRN_DP_OP", "sum") == "skip":      # tooling: time the step without the collective
        return
    if os.environ.get("PRN_DP_OP", "sum") == "avg":
        dist.all_reduce(t, op=dist.ReduceOp.AVG)
    else:
        dist.all_reduce(t)
        t.mul_(1.0 / world)


class GraphedStep:
    """Forward and backward of one training step captured as two CUDA graphs sharing a memory pool (the ~1300 launches
    of a step cost more host time than GPU time otherwise).  Weight packing is captured too, so that replays see the
    optimizer's in-place parameter updates; parameters must keep their storage (true for torch optimizers)."""

    def __init__(self, eng, net, x, pack_in_graph=True, optimizer=None, world=1, buckets=None, dp_mode=None, flat_grads=None):
        """optimizer: a planerecnet_b200.optim.FusedAdam over net's parameters -> `optimizer_step()` replays the update
        from a third graph (after the gradient all-reduce when world > 1).  buckets: capture the backward as one graph per
        gradient bucket (default: only when world > 1; True on one GPU exercises the same path without the collective).
        dp_mode (world > 1; default from PRN_DP_MODE, else 'after'):
          'after'        ONE backward graph that ends with the multi-tensor copy of all gradients into the persistent flat fp32
                         buffer, then one NCCL all-reduce (avg) over it (exactly one all-reduce per step);
          'p2p'          backward in three bucket graphs, per-bucket mean all-reduce through peer memory with the copy engines
                         (utils.dist.PeerAllReduce) on a communication stream, started as each bucket's graph has been enqueued;
          'nccl_overlap' the same schedule with NCCL all-reduces.
        flat_grads: end the (single) backward graph with the gather of every parameter gradient into ONE persistent flat fp32
        buffer (ops.CopyMulti: one launch) — default when world > 1 (the all-reduce operand); the autograd boundary asks for
        it too and hands autograd views of one clone of that buffer instead of ~400 per-parameter copies.
        Measured at N = 2 (R101 bs 8, tools/dp_check.py, profiles/r02_dp_check_2gpu.txt): cutting the backward into bucket graphs
        costs more than the overlap wins — every graph boundary joins the side-stream weight gradients, which lag the main chain —
        so 'after' is the default."""
        import os
        self.dp_mode = dp_mode or os.environ.get("PRN_DP_MODE", "after")
        assert self.dp_mode in ("p2p", "nccl_overlap", "after")
        if buckets is None:
            buckets = world > 1 and self.dp_mode != "after"
        self.eng, self.net = eng, net
        self.sx = torch.empty_like(x)
        self.sx.copy_(x)
        self.bns = [m for m in net.modules() if isinstance(m, nn.BatchNorm2d) and m.training and m.track_running_stats]
        # the eager warm-up below is not a training step: it must leave the BatchNorm buffers (running statistics through
        # raw pointers, num_batches_tracked) exactly as it found them, or every newly captured graph would apply its
        # first batch twice
        bn_bufs = [b for m in self.bns for b in (m.running_mean, m.running_var, m.num_batches_tracked)]
        bn_saved = [b.detach().clone() for b in bn_bufs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                 # eager warm-up: function attributes, allocator, first packs
            outs = eng.forward_train(net, self.sx)
            eng.seed_output_grads(torch.zeros_like(outs[0]), [torch.zeros_like(c) for c in outs[1]],
                                  [torch.zeros_like(k) for k in outs[2]], torch.zeros_like(outs[3]))
            eng._fin_single = set()
            warm_ids = set(eng.backward().keys())
            defer_ids = set(eng._fin_single)          # conv weights whose gradient is one accumulator -> one deferred unpack
            with torch.no_grad():
                for b, sv in zip(bn_bufs, bn_saved):
                    b.copy_(sv)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        eng.reset()                                      # allocates the large accumulator arena the warm-up sized (never inside a capture)
        if pack_in_graph:
            eng.repack_all(invalidate_others=False)      # uploads the pack tables (not allowed inside a capture); same values
            torch.cuda.synchronize()
        self.g_fwd = torch.cuda.CUDAGraph()
        n0 = eng.launches
        with _no_gc(), torch.cuda.graph(self.g_fwd):
            if pack_in_graph and not eng.repack_all():
                eng._packed.clear()
            self.outs = eng.forward_train(net, self.sx)
        self.fwd_launches = eng.launches - n0
        m, cs, ks, d = self.outs
        self.cots = (torch.zeros_like(m), [torch.zeros_like(c) for c in cs], [torch.zeros_like(k) for k in ks], torch.zeros_like(d))
        self.optimizer, self.world = optimizer, world
        self.g_opt = self.g_flat = None
        self.flat = None
        self._reduced = False
        self.g_bwd_seg = None
        if not buckets:
            params = [p for p in net.parameters() if p.requires_grad and id(p) in warm_ids]
            use_flat = world > 1 if flat_grads is None else bool(flat_grads)
            gather = None
            if use_flat:
                self.flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device="cuda")
                views, off = {}, 0
                for p in params:
                    views[id(p)] = self.flat[off:off + p.numel()].view(p.shape)
                    off += p.numel()
                self.flat_views, self.bucket_slices = views, [(0, off)]
                self.flat_params = params
                # deferred weight gradients are unpacked straight into their flat views: the gather only moves the others
                rest = [p for p in params if id(p) not in defer_ids]
                gather = ops.CopyMulti([views[id(p)] for p in rest]) if rest else None
            um = ops.UnpackMulti()
            self.g_bwd = torch.cuda.CUDAGraph()
            n0 = eng.launches
            with _no_gc(), torch.cuda.graph(self.g_bwd, pool=self.g_fwd.pool()):
                eng.seed_output_grads(*self.cots)
                eng._defer = dict(ids=defer_ids, views=views if use_flat else None, um=um, dst_ids=set())
                try:
                    self.grads = eng.backward()
                finally:
                    eng._defer = None
                if use_flat and gather is not None:
                    srcs = gather.normalise([self.grads[id(p)].reshape(p.shape) for p in rest])
                    gather.run()
                    eng.launches += 1
            if um.ran:
                um.flush()                            # table contents of the deferred unpack: read by the launch at replay time
            self._unpack = um
            if use_flat and gather is not None:
                gather.set_sources(srcs)              # table contents: read by the launch at replay time
                self._gather = gather
            self.bwd_launches = eng.launches - n0
            params = [p for p in params if id(p) in self.grads]
            src = views if use_flat else self.grads
        else:
            # Data parallel: the backward is captured as N_BUCKETS graphs cut where a gradient bucket becomes final
            # (FPN + heads + decoder | backbone layers 2-3 | stem + layers 0-1).  Each graph ends with the multi-tensor copy of
            # its bucket into a contiguous slice of ONE persistent flat fp32 buffer; `backward` replays the graphs and, after
            # each, starts that slice's NCCL all-reduce on a communication stream, so only the last (smallest) bucket's
            # all-reduce is exposed — the others hide under the remaining backward.
            bucket = eng.bucket_of_params(net)
            params = [p for p in net.parameters() if p.requires_grad and id(p) in warm_ids]
            params.sort(key=lambda p: bucket[id(p)])                      # stable: bucket-major, module order inside
            self.flat = torch.empty(sum(p.numel() for p in params), dtype=torch.float32, device="cuda")
            views, off, self.bucket_slices = {}, 0, []
            for b in range(eng.N_BUCKETS):
                lo = off
                for p in params:
                    if bucket[id(p)] == b:
                        views[id(p)] = self.flat[off:off + p.numel()].view(p.shape)
                        off += p.numel()
                self.bucket_slices.append((lo, off))
            self.flat_views = views
            self.g_bwd_seg = [torch.cuda.CUDAGraph() for _ in range(eng.N_BUCKETS)]
            n0 = eng.launches
            for b in range(eng.N_BUCKETS):
                torch.cuda.synchronize()
                with _no_gc(), torch.cuda.graph(self.g_bwd_seg[b], pool=self.g_fwd.pool()):
                    if b == 0:
                        eng.seed_output_grads(*self.cots)
                    pg = eng.backward_segment(b)
                    mine = [p for p in params if bucket[id(p)] == b]
                    missing = [p for p in mine if id(p) not in pg]
                    assert not missing, f"bucket {b}: {len(missing)} parameter gradients are not final after their tape segment"
                    torch._foreach_copy_([views[id(p)] for p in mine], [pg[id(p)].reshape(p.shape) for p in mine])
                    if b == eng.N_BUCKETS - 1:
                        self.grads = pg
            self.bwd_launches = eng.launches - n0
            self.comm = torch.cuda.Stream()
            self.peer = None
            if world > 1 and self.dp_mode == "p2p":
                from .utils.dist import PeerAllReduce
                torch.cuda.synchronize()
                self.peer = PeerAllReduce(self.flat, self.bucket_slices)
            src = views
        if optimizer is not None:
            optimizer.prepare(src)
            torch.cuda.synchronize()
            self.g_opt = torch.cuda.CUDAGraph()
            with _no_gc(), torch.cuda.graph(self.g_opt, pool=self.g_fwd.pool()):
                optimizer.step(src)
            # the capture executed nothing, but step() bumped versions / the warm-up above ran one real forward+backward:
            # parameters are untouched; optimizer state starts at step 0

    def allreduce_grads(self):
        """Gradients of the last backward averaged over the ranks.  With world > 1 the per-bucket NCCL all-reduces (ReduceOp.AVG
        over slices of the persistent flat fp32 buffer) were started by `backward` as each bucket became final: this only makes
        the caller's stream wait for the communication stream.  Returns {id(param): averaged view}."""
        if self.g_bwd_seg is None:
            self._reduced = self.flat is not None
            return self.flat_views if self.flat is not None else self.grads
        torch.cuda.current_stream().wait_stream(self.comm)
        self._reduced = True
        return self.flat_views

    def allreduce_desc(self):
        if self.flat is None:
            return None
        mb = [f"{(hi - lo) * 4 / 2 ** 20:.0f}" for lo, hi in self.bucket_slices]
        how = {"p2p": "mean all-reduce through peer memory: copy-engine pulls over NVLink (reduce-scatter + all-gather), one small sum "
                      "kernel and 4-byte NCCL all-reduces as inter-rank barriers",
               "nccl_overlap": "NCCL all-reduce (avg)", "after": "ONE NCCL all-reduce (avg) over the whole buffer after the backward"}[self.dp_mode]
        return (f"{how}; one persistent flat fp32 gradient buffer in {len(mb)} buckets ({' + '.join(mb)} MiB, reverse execution order)"
                + ("" if self.dp_mode == "after" else ", each bucket started on a communication stream as its backward graph has been enqueued"))

    def optimizer_step(self):
        """(all-reduce the gradients of the last backward over the ranks and) apply the optimizer, from graphs."""
        if self.g_bwd_seg is not None and not self._reduced:
            self.allreduce_grads()          # bucket modes: make this stream wait for the communication stream
        self._reduced = False
        self.g_opt.replay()
        for p in self.net.parameters():
            torch.autograd.graph.increment_version(p)

    def forward(self, x):
        self.sx.copy_(x, non_blocking=True)
        self.g_fwd.replay()
        self.eng.launches += self.fwd_launches
        for bn in self.bns:      # the replay updated the running statistics through raw pointers
            torch.autograd.graph.increment_version(bn.running_mean)
            torch.autograd.graph.increment_version(bn.running_var)
        return self.outs

    def backward(self, d_mask, d_cates, d_kerns, d_depth):
        def put(dst, src):
            if src is None:
                dst.zero_()
            else:
                dst.copy_(src, non_blocking=True)

        put(self.cots[0], d_mask)
        for dst, src in zip(self.cots[1], d_cates):
            put(dst, src)
        for dst, src in zip(self.cots[2], d_kerns):
            put(dst, src)
        put(self.cots[3], d_depth)
        self.eng.launches += self.bwd_launches
        self._reduced = False
        if self.g_bwd_seg is None:
            self.g_bwd.replay()
            if self.world > 1:
                import torch.distributed as dist
                if dist.is_initialized():
                    _allreduce_mean(self.flat, self.world)
            return self.grads
        import torch.distributed as dist
        main = torch.cuda.current_stream()
        live = self.world > 1 and dist.is_initialized()
        self.comm.wait_stream(main)          # the previous step's consumers of `flat` (optimizer) are done before it is rewritten
        for b, g in enumerate(self.g_bwd_seg):
            g.replay()
            lo, hi = self.bucket_slices[b]
            if not live or self.dp_mode == "after":
                continue
            ev = torch.cuda.Event()
            ev.record(main)
            self.comm.wait_event(ev)
            with torch.cuda.stream(self.comm):
                if self.peer is not None:
                    self.peer.reduce(b)
                elif hi > lo:
                    _allreduce_mean(self.flat[lo:hi], self.world)
        if live:
            if self.dp_mode == "after":
                _allreduce_mean(self.flat, self.world)                    # main stream: nothing left to overlap with
            elif self.peer is not None:
                with torch.cuda.stream(self.comm):
                    self.peer.finish()
        return self.grads


class _DenseTrainFn(torch.autograd.Function):
    """Autograd boundary of the training branch: one node whose backward runs the sm_100a tape."""

    @staticmethod
    def forward(ctx, net, x, *params):
        eng = net.train_engine
        ctx.step = None
        if getattr(net, "use_train_graph", False):
            bns = getattr(net, "_bn_modules_cache", None)
            if bns is None:        # walking the module tree costs 1.7 ms per call (R101: 17 000 named_modules visits) with the GPU idle
                bns = net._bn_modules_cache = [m for m in net.modules() if isinstance(m, nn.BatchNorm2d)]
            key = (id(net), tuple(x.shape), tuple(p.requires_grad for p in params), tuple(m.training for m in bns))
            step = eng._graphs_t.get(key)
            if step is None:
                step = eng._graphs_t[key] = GraphedStep(eng, net, x, flat_grads=True)
            mask, cates, kerns, depth = step.forward(x)
            # the graph's static buffers are overwritten by the next step: hand out copies
            mask, cates, kerns, depth = mask.clone(), [c.clone() for c in cates], [k.clone() for k in kerns], depth.clone()
            ctx.step = step
        else:
            mask, cates, kerns, depth = eng.forward_train(net, x)
        ctx.eng, ctx.params = eng, params
        ctx.nl = len(cates)
        return (mask, *cates, *kerns, depth)

    @staticmethod
    def backward(ctx, *cots):
        eng, nl = ctx.eng, ctx.nl
        args = (cots[0], cots[1:1 + nl], cots[1 + nl:1 + 2 * nl], cots[1 + 2 * nl])
        if ctx.step is not None:
            g = ctx.step.backward(*args)
        else:
            eng.seed_output_grads(*args)
            g = eng.backward()
        # copies: the accumulators (and, when graphed, the static buffers) are reused by the next step.  All parameter gradients
        # go into ONE fresh flat buffer with a multi-tensor copy (two launches instead of one per parameter: ~450); autograd
        # receives views of it.
        out = [None] * len(ctx.params)
        step = ctx.step
        if step is not None and step.flat is not None and step.g_bwd_seg is None:
            # graphed: the backward graph has gathered every gradient into the step's persistent flat buffer; autograd gets views
            # of ONE clone of it (the buffer itself is rewritten by the next replay)
            flat = step.flat.clone()
            if eng.grad_scale != 1.0:
                flat.mul_(1.0 / eng.grad_scale)
            pos = {id(p): i for i, p in enumerate(ctx.params)}
            off = 0
            for p in step.flat_params:
                i = pos.get(id(p))
                if i is not None and p.requires_grad:
                    assert p.dtype == torch.float32
                    out[i] = flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
            return (None, None, *out)
        idx, srcs = [], []
        for i, p in enumerate(ctx.params):
            gp = g.get(id(p))
            if gp is not None and p.requires_grad:
                assert p.dtype == torch.float32
                idx.append(i)
                srcs.append(gp.reshape(p.shape))
        if idx:
            sizes = [ctx.params[i].numel() for i in idx]
            flat = torch.empty(sum(sizes), dtype=torch.float32, device=srcs[0].device)
            views = [v.view(ctx.params[i].shape) for v, i in zip(flat.split(sizes), idx)]
            torch._foreach_copy_(views, srcs)
            if eng.grad_scale != 1.0:
                flat.mul_(1.0 / eng.grad_scale)
            for i, v in zip(idx, views):
                out[i] = v
        return (None, None, *out)


def forward_train_autograd(net, x):
    # net.parameters() walks the module tree (R101: ~1 ms per call, with the GPU idle at the start of a step): the list is cached
    # on the module and dropped by PlaneRecNet._apply / load_state_dict (device moves, checkpoint loads)
    params = getattr(net, "_params_cache", None)
    if params is None:
        params = net._params_cache = [p for p in net.parameters()]
    outs = _DenseTrainFn.apply(net, x, *params)
    nl = (len(outs) - 2) // 2
    return outs[0], list(outs[1:1 + nl]), list(outs[1 + nl:1 + 2 * nl]), outs[-1]
