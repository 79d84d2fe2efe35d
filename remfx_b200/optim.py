"""L5 -- the optimiser half of the reference's training step on the GPU.

Reference: `RemFX.configure_optimizers` (remfx/models.py:185-206) = torch.optim.AdamW(lr 1e-4, betas (0.95, 0.999),
eps 1e-6, weight_decay 1e-3) + MultiStepLR([0.8, 0.95] * max_steps, gamma 0.1) stepped every batch, with Lightning's
`gradient_clip_val: 10.0` (cfg/config.yaml:119) and DDP gradient averaging.

Here every parameter / gradient / moment lives in ONE flat fp32 bucket (`FlatBucket`), so a step is
    [all-reduce(bucket) over NCCL]  ->  rfx_grad_sumsq  ->  rfx_adamw_step   (csrc/optim.cu)
i.e. one collective and two streaming kernels for the whole model; averaging (1 / world) and the clip coefficient are
folded into the update kernel.  `FusedAdamW` is a torch.optim.Optimizer, so the reference's MultiStepLR (and Lightning)
drive it unchanged.  The kernels have no CPU fallback: `step()` on CPU tensors raises RfxError.
"""
from __future__ import annotations

import weakref
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib

_ALIGN = 64  # elements (256 B): every parameter view starts on a 256-byte boundary of the bucket


class FlatBucket:
    """Flat fp32 storage for a parameter list.  After construction every `p.data` (and `p.grad`) is a view into
    `self.param` (`self.grad`); padding between views stays zero so whole-bucket kernels are safe."""

    def __init__(self, params: Iterable[Tensor]):
        self.params: List[Tensor] = [p for p in params]
        if not self.params:
            raise ValueError("FlatBucket needs at least one parameter")
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError("FlatBucket holds fp32 parameters only (the reference trains in fp32, cfg/config.yaml:112)")
            if p.device != dev:
                raise ValueError("all parameters must live on one device")
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.numel = max(off, _ALIGN)
        self.param = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                view = self.param[o:o + p.numel()].view(p.shape)
                view.copy_(p.data)
                old_grad = p.grad
                p.data = view
                p.grad = self.grad[o:o + p.numel()].view(p.shape)
                if old_grad is not None:
                    p.grad.copy_(old_grad)
        # gradient sink (alloc_param_grads below): a network's hand-written backward can ask for its parameter gradients as views
        # of ONE zeroed flat buffer laid out like this bucket; collect_grads() then adopts that buffer instead of copying
        self._incoming: Optional[Tensor] = None
        ref = weakref.ref(self)
        for i, p in enumerate(self.params):
            p._rfx_sink = (ref, i)

    def sink_alloc(self) -> Tensor:
        """A zeroed flat gradient buffer with this bucket's layout (one memset for the whole model); remembered until the
        next collect_grads()."""
        self._incoming = torch.zeros(self.numel, dtype=torch.float32, device=self.param.device)
        return self._incoming

    def grad_view(self, i: int) -> Tensor:
        p, o = self.params[i], self.offsets[i]
        return self.grad[o:o + p.numel()].view(p.shape)

    def param_view(self, i: int) -> Tensor:
        p, o = self.params[i], self.offsets[i]
        return self.param[o:o + p.numel()].view(p.shape)

    def collect_grads(self) -> List[int]:
        """Make `self.grad` hold every parameter's gradient: a no-op while `.grad` still aliases the bucket (autograd
        accumulates in place), a copy for gradients that were re-created (e.g. after zero_grad(set_to_none=True)).
        Returns the indices of parameters that have NO gradient this step (torch.optim skips those entirely)."""
        missing: List[int] = []
        inc, self._incoming = self._incoming, None
        if inc is not None and inc.numel() == self.numel and inc.device == self.grad.device:
            # fast path: every gradient autograd stored is (still) the view of the sink buffer at this bucket's offset -> the
            # buffer BECOMES the bucket's gradient storage: no per-parameter accumulate / copy kernels at all
            base = inc.data_ptr()
            ok = True
            for i, p in enumerate(self.params):
                g = p.grad
                if g is None:
                    missing.append(i)
                elif g.data_ptr() != base + 4 * self.offsets[i] or g.numel() != p.numel() or not g.is_contiguous():
                    ok = False
                    break
            if ok:
                self.grad = inc
                with torch.no_grad():
                    for i in missing:
                        self.grad_view(i).zero_()
                return missing
            missing = []
        with torch.no_grad():
            for i, p in enumerate(self.params):
                view = self.grad_view(i)
                g = p.grad
                if g is None:
                    view.zero_()
                    missing.append(i)
                elif g.data_ptr() != view.data_ptr():
                    view.copy_(g)
                    p.grad = view
        return missing

    def rebind_params(self) -> int:
        """Re-point every parameter whose storage left the bucket (module.to / .float() / load with assign=True after the
        optimiser was built) at its bucket view, copying its current value in; returns how many were re-bound.  Without this
        the update kernel would keep training the bucket while the live parameters silently stopped moving."""
        n = 0
        with torch.no_grad():
            for i, p in enumerate(self.params):
                view = self.param_view(i)
                if p.data_ptr() == view.data_ptr():
                    continue
                if p.device != self.param.device or p.dtype != torch.float32 or p.numel() != view.numel():
                    raise RuntimeError(f"FlatBucket: parameter {i} moved to {p.device}/{p.dtype} after the optimiser was built; "
                                       "rebuild the optimiser")
                view.copy_(p.data.reshape(view.shape))
                p.data = view
                n += 1
        return n

    def zero_grad(self, set_to_none: bool = False) -> None:
        """Default: gradients stay zeroed views of the bucket (autograd accumulates in place).  set_to_none=True (torch's own
        default): every `.grad` is dropped, so that autograd STORES the next gradients instead of adding them -- with the
        gradient sink that makes a step free of per-parameter kernels."""
        if set_to_none:
            for p in self.params:
                p.grad = None
            return
        self.grad.zero_()
        for i, p in enumerate(self.params):
            if p.grad is None or p.grad.data_ptr() != self.grad_view(i).data_ptr():
                p.grad = self.grad_view(i)


def alloc_param_grads(params: Sequence[Tensor]) -> Tuple[List[Tensor], bool]:
    """Gradient buffers for a hand-written backward.  When every trainable parameter belongs to one live FlatBucket and holds no
    gradient yet (zero_grad(set_to_none=True)), they are views of one zeroed flat buffer in the bucket's layout (second value
    True: the buffers are already zero); otherwise plain uninitialised tensors (False)."""
    bucket = None
    usable = True
    for p in params:
        if not p.requires_grad:
            continue
        sink = getattr(p, "_rfx_sink", None)
        b = sink[0]() if sink is not None else None
        if b is None or (bucket is not None and b is not bucket) or p.grad is not None or b.params[sink[1]] is not p:
            usable = False
            break
        bucket = b
    if not usable or bucket is None:
        return [torch.empty_like(p, memory_format=torch.contiguous_format) for p in params], False
    flat = bucket.sink_alloc()
    grads = []
    for p in params:
        if p.requires_grad:
            o = bucket.offsets[p._rfx_sink[1]]
            grads.append(flat[o:o + p.numel()].view(p.shape))
        else:
            grads.append(torch.zeros_like(p, memory_format=torch.contiguous_format))
    return grads, True


def sync_grads(bucket_grad: Tensor, group=None) -> float:
    """SUM all-reduce of the flat gradient bucket across data-parallel ranks (one collective for the whole model);
    returns the scale (1 / world) the update kernel folds in.  gloo on CPU in the tests, NCCL over NVLink on the box."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(bucket_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


def multistep_lr(step: int, max_steps: int, base_lr: float = 1e-4, gamma: float = 0.1) -> float:
    """Learning rate in force for the optimiser update number `step` (0-based count of completed scheduler steps), as
    MultiStepLR([0.8 * max_steps, 0.95 * max_steps], gamma) stepped every batch yields (remfx/models.py:192-196)."""
    m1, m2 = 0.8 * max_steps, 0.95 * max_steps
    return base_lr * (gamma ** (int(step >= m1) + int(step >= m2)))


class FusedAdamW(torch.optim.Optimizer):
    """Drop-in for the reference's `torch.optim.AdamW(list(model.parameters()), ...)` with gradient averaging across
    ranks and clip-by-global-norm folded in.  One parameter group (what the reference builds)."""

    def __init__(self, params, lr: float = 1e-4, betas=(0.95, 0.999), eps: float = 1e-6, weight_decay: float = 1e-3,
                 max_grad_norm: Optional[float] = 10.0, process_group=None):
        params = list(params)
        if params and isinstance(params[0], dict):
            raise ValueError("FusedAdamW takes a flat parameter list (one group), like the reference")
        defaults = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.max_grad_norm = max_grad_norm
        self.process_group = process_group
        # torch.optim.AdamW never touches parameters without a gradient: frozen ones (requires_grad=False) stay out of the
        # bucket altogether; trainable ones that happen to have no gradient in a step are restored after the kernel (step()).
        self.bucket = FlatBucket([p for p in self.param_groups[0]["params"] if p.requires_grad])
        dev = self.bucket.param.device
        self.exp_avg = torch.zeros_like(self.bucket.param)
        self.exp_avg_sq = torch.zeros_like(self.bucket.param)
        self.step_count = 0
        self._ws = torch.zeros(32, dtype=torch.float64, device=dev)      # [0] = sum of squares
        self.total_norm = torch.zeros(1, dtype=torch.float32, device=dev)  # pre-clip global norm of the last step

    def zero_grad(self, set_to_none: bool = False) -> None:
        # default: gradients stay views of the bucket; set_to_none=True: dropped, the gradient sink stores the next ones
        self.bucket.zero_grad(set_to_none)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        b = self.bucket
        _lib.require_device(b.param)
        b.rebind_params()
        missing = b.collect_grads()
        kept = [(i, b.param_view(i).clone(), self._moment_views(i)) for i in missing]
        kept = [(i, pv, (m.clone(), v.clone())) for i, pv, (m, v) in kept]
        ev = self._timing_events()
        if ev:
            ev[0].record()
        scale = sync_grads(b.grad, self.process_group)   # the ONE collective of a data-parallel step
        if ev:
            ev[1].record()
        g = self.param_groups[0]
        self.step_count += 1
        L = _lib.lib()
        clip = float(self.max_grad_norm) if self.max_grad_norm else 0.0
        with torch.cuda.device(b.param.device):
            s = _lib.cur_stream()
            if clip > 0:
                _lib.check(L.rfx_grad_sumsq(b.grad.data_ptr(), b.numel, self._ws.data_ptr(), 0, s), "rfx_grad_sumsq")
            _lib.check(L.rfx_adamw_step(b.param.data_ptr(), b.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                        b.numel, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                                        float(g["weight_decay"]), self.step_count, float(scale), clip, self._ws.data_ptr(),
                                        self.total_norm.data_ptr(), s), "rfx_adamw_step")
        # the kernel wrote the parameters through raw pointers: tell autograd (and the model handles, whose cached
        # packed weights are keyed on `_version`) that they changed
        if ev:
            ev[2].record()
        for i, pv, (m0, v0) in kept:  # no gradient this step: no decay, no moment update (what torch.optim.AdamW does)
            b.param_view(i).copy_(pv)
            m, v = self._moment_views(i)
            m.copy_(m0)
            v.copy_(v0)
        torch.autograd.graph.increment_version(b.params)
        return loss

    # ---- optional per-step device timing of the collective and the update kernels (bench.py's train leg)
    def set_timing(self, on: bool) -> None:
        self._timing = bool(on)
        self._timed = []

    def _timing_events(self):
        if not getattr(self, "_timing", False):
            return None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        self._timed.append(ev)
        return ev

    def timing_ms(self):
        """[(all_reduce_ms, clip+adamw_ms)] of every timed step; the device must be synchronised first."""
        return [(e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])) for e in getattr(self, "_timed", [])]

    def _group_indices(self) -> List[int]:
        """Position of every bucketed (trainable) parameter in the optimiser's parameter group (torch's state_dict index)."""
        pos = {id(p): k for k, p in enumerate(self.param_groups[0]["params"])}
        return [pos[id(p)] for p in self.bucket.params]

    def _moment_views(self, i: int):
        p, o = self.bucket.params[i], self.bucket.offsets[i]
        return self.exp_avg[o:o + p.numel()].view(p.shape), self.exp_avg_sq[o:o + p.numel()].view(p.shape)

    # state_dict in torch.optim.AdamW's layout (per-parameter exp_avg / exp_avg_sq / step) so checkpoints interchange
    def state_dict(self):
        b = self.bucket
        state = {}
        gidx = self._group_indices()
        for i, (p, o) in enumerate(zip(b.params, b.offsets)):
            state[gidx[i]] = {
                "step": torch.tensor(float(self.step_count)),
                "exp_avg": self.exp_avg[o:o + p.numel()].view(p.shape).clone(),
                "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view(p.shape).clone(),
            }
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        g["params"] = list(range(len(self.param_groups[0]["params"])))
        return {"state": state, "param_groups": [g]}

    def load_state_dict(self, sd) -> None:
        b = self.bucket
        for k, v in sd["param_groups"][0].items():
            if k != "params":
                self.param_groups[0][k] = v
        with torch.no_grad():
            gidx = self._group_indices()
            for i, (p, o) in enumerate(zip(b.params, b.offsets)):
                st = sd["state"].get(gidx[i])
                if st is None:
                    continue
                self.exp_avg[o:o + p.numel()].view(p.shape).copy_(st["exp_avg"])
                self.exp_avg_sq[o:o + p.numel()].view(p.shape).copy_(st["exp_avg_sq"])
                self.step_count = int(float(st["step"]))


def configure_optimizers(module: torch.nn.Module, max_steps: int, lr: float = 1e-4, lr_beta1: float = 0.95, lr_beta2: float = 0.999,
                         lr_eps: float = 1e-6, lr_weight_decay: float = 1e-3, gradient_clip_val: float = 10.0, process_group=None):
    """Same return structure as `RemFX.configure_optimizers` (remfx/models.py:185-206)."""
    optimizer = FusedAdamW(list(module.parameters()), lr=lr, betas=(lr_beta1, lr_beta2), eps=lr_eps, weight_decay=lr_weight_decay,
                           max_grad_norm=gradient_clip_val, process_group=process_group)
    lr_scheduler = torch.optim.lr_scheduler.MultiStepLR(optimizer, [0.8 * max_steps, 0.95 * max_steps], gamma=0.1)
    return {"optimizer": optimizer,
            "lr_scheduler": {"scheduler": lr_scheduler, "monitor": "val_loss", "interval": "step", "frequency": 1}}
