"""Adam for the training step (train.py:251-256: `optim.Adam` over five parameter groups with their own learning rates):
one multi-tensor launch of `prn_adam_multi` over every parameter, with the step counter in device memory so that the
update can be captured into the step's CUDA graphs.  Same update rule and state names as torch.optim.Adam.

Differences from torch.optim.Adam: ONE step counter for all parameters — every parameter has to receive a gradient in
every step (true for the PlaneRecNet training step; torch keeps a counter per parameter, so a parameter that skips steps would get a
different bias correction there); `skip_nonfinite` (default: on when `grad_scale != 1`, i.e. loss-scaled f16 training) checks all
gradients first and skips the whole update — parameters, moments and the step counter — when one is inf / NaN, like
torch.cuda.amp.GradScaler.step; `found_inf` (device int32 tensor) tells the caller to lower its scale.  The pointer tables are
rebuilt (a few host-to-device copies) whenever a gradient tensor moves: keep gradients in persistent buffers on the hot path
(`GraphedStep.optimizer_step()` does: the flat gradient buffer never moves)."""
import ctypes as C

import torch

from . import _lib as L

_CHUNK = 1 << 16


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_scale=1.0, skip_nonfinite=None):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        self.param_groups = [{"params": list(g["params"]), "lr": g.get("lr", lr)} for g in groups]
        self.betas, self.eps, self.grad_scale = betas, eps, grad_scale
        self.skip_nonfinite = (grad_scale != 1.0) if skip_nonfinite is None else bool(skip_nonfinite)
        self.found_inf = None            # device int32 [1]: 1 after a skipped step (skip_nonfinite)
        self.state = {}
        self._state3 = None
        self._key = None
        self._lr_key = None
        self._tabs = None

    def _params(self):
        return [(p, g) for g in self.param_groups for p in g["params"] if p.requires_grad]

    def _build(self, grads):
        items = [(p, grp, grads[id(p)]) for p, grp in self._params() if id(p) in grads and grads[id(p)] is not None]
        key = tuple((p.data_ptr(), g.data_ptr(), tuple(g.stride())) for p, _, g in items)
        lr_key = tuple(float(grp["lr"]) for _, grp, _ in items)
        if key == self._key:
            if lr_key != self._lr_key:      # set_lr (train.py:415-420): refresh the table in place, captured graphs keep seeing it
                self._tabs[2].copy_(torch.tensor(lr_key, dtype=torch.float32))
                self._lr_key = lr_key
            return
        self._lr_key = lr_key
        dev = items[0][0].device
        table, numel, lrs, chunks = [], [], [], []
        for t, (p, grp, g) in enumerate(items):
            assert p.dtype == torch.float32 and p.is_contiguous() and g.dtype == torch.float32 and g.numel() == p.numel()
            if g.is_contiguous():
                gs = 1
            else:
                assert g.dim() == 1, "gradient views must be contiguous or 1-D strided"
                gs = g.stride(0)
            st = self.state.setdefault(id(p), {})
            if not st:
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            table += [p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), gs]
            numel.append(p.numel())
            lrs.append(grp["lr"])
            chunks += [(t, c) for c in range((p.numel() + _CHUNK - 1) // _CHUNK)]
        self._tabs = (torch.tensor(table, dtype=torch.int64, device=dev), torch.tensor(numel, dtype=torch.int64, device=dev),
                      torch.tensor(lrs, dtype=torch.float32, device=dev),
                      torch.tensor(chunks, dtype=torch.int32, device=dev).contiguous(), len(chunks), [g for _, _, g in items])
        if self._state3 is None:
            self._state3 = torch.zeros(3, dtype=torch.float32, device=dev)
        if self.found_inf is None:
            self.found_inf = torch.zeros(1, dtype=torch.int32, device=dev)
        self._key = key

    def prepare(self, grads):
        """Build the pointer tables for {id(param): grad} (host-to-device copies): call before graph capture."""
        self._build(grads)

    def step(self, grads=None):
        """grads: {id(param): fp32 tensor} (e.g. TrainEngine.backward()'s result); default: the parameters' .grad."""
        if grads is None:
            grads = {id(p): p.grad for p, _ in self._params()}
        self._build(grads)
        table, numel, lrs, chunks, n_chunks, _keep = self._tabs
        if self.skip_nonfinite:
            L.check(L.lib().prn_adam_multi_checked(C.c_void_p(table.data_ptr()), C.c_void_p(numel.data_ptr()), C.c_void_p(lrs.data_ptr()),
                                                   C.c_void_p(chunks.data_ptr()), n_chunks, C.c_void_p(self._state3.data_ptr()),
                                                   C.c_float(self.betas[0]), C.c_float(self.betas[1]), C.c_float(self.eps),
                                                   C.c_float(self.grad_scale), C.c_void_p(self.found_inf.data_ptr()),
                                                   L.current_stream()), "prn_adam_multi_checked")
        else:
            L.check(L.lib().prn_adam_multi(C.c_void_p(table.data_ptr()), C.c_void_p(numel.data_ptr()), C.c_void_p(lrs.data_ptr()),
                                           C.c_void_p(chunks.data_ptr()), n_chunks, C.c_void_p(self._state3.data_ptr()),
                                           C.c_float(self.betas[0]), C.c_float(self.betas[1]), C.c_float(self.eps),
                                           C.c_float(self.grad_scale), L.current_stream()), "prn_adam_multi")
        for p, _ in self._params():       # raw-pointer update: bump the version counters the weight caches key on
            torch.autograd.graph.increment_version(p)

    def zero_grad(self, set_to_none=True):
        for p, _ in self._params():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()
