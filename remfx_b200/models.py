"""Drop-in network wrappers: the reference's plug-in boundary (`cfg/model/*.yaml` `_target_`s).

Same constructor kwargs, `forward((x, target)) -> (loss, output)` / `sample(x) -> output`
signatures and `state_dict` key layout as `remfx.models.{OpenUnmixModel, TCNModel, DemucsModel}`
(remfx/models.py:259-390), so `cfg/exp/*` can select them by overriding `model.network._target_`.
The torch.nn sub-modules below are *parameter containers only* (they give the reference's
state_dict keys and default initialisation); their `forward` is never called -- all math runs in
the hand-written sm_100a kernels behind the C ABI (include/remfx_b200.h).  There is no CPU or
eager fallback: calling these modules with CPU tensors raises.

Parity is defined in eval mode: BatchNorm uses running statistics
and dropout is off, which is how the reference's chain/test scripts are meant to run the models.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor, nn

from . import _lib


class _OpenUnmixParams(nn.Module):
    """Parameter layout of umx/openunmix/model.py:33-98 (OpenUnmix.__init__), bidirectional."""

    def __init__(self, nb_bins: int, nb_channels: int = 1, hidden_size: int = 512, nb_layers: int = 3):
        super().__init__()
        self.nb_bins = nb_bins
        self.hidden_size = hidden_size
        self.nb_layers = nb_layers
        self.fc1 = nn.Linear(nb_bins * nb_channels, hidden_size, bias=False)
        self.bn1 = nn.BatchNorm1d(hidden_size)
        self.lstm = nn.LSTM(input_size=hidden_size, hidden_size=hidden_size // 2, num_layers=nb_layers, bidirectional=True,
                            batch_first=False, dropout=0.4 if nb_layers > 1 else 0)
        self.fc2 = nn.Linear(hidden_size * 2, hidden_size, bias=False)
        self.bn2 = nn.BatchNorm1d(hidden_size)
        self.fc3 = nn.Linear(hidden_size, nb_bins * nb_channels, bias=False)
        self.bn3 = nn.BatchNorm1d(nb_bins * nb_channels)
        self.input_mean = nn.Parameter(torch.zeros(nb_bins))
        self.input_scale = nn.Parameter(torch.ones(nb_bins))
        self.output_scale = nn.Parameter(torch.ones(nb_bins))
        self.output_mean = nn.Parameter(torch.ones(nb_bins))

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container; the computation lives in libremfx_b200.so")


class _WindowHolder(nn.Module):
    def __init__(self, n_fft: int):
        super().__init__()
        self.window = nn.Parameter(torch.hann_window(n_fft), requires_grad=False)


class _SeparatorParams(nn.Module):
    """State layout of umx/openunmix/model.py:197-240 (Separator.__init__)."""

    def __init__(self, target_models, sample_rate: float, n_fft: int):
        super().__init__()
        self.stft = _WindowHolder(n_fft)
        self.istft = _WindowHolder(n_fft)
        self.target_models = nn.ModuleDict(target_models)
        self.register_buffer("sample_rate", torch.as_tensor(sample_rate))  # dtype follows the argument, as in the reference (int -> int64)


class OpenUnmixModel(nn.Module):
    """B200-native drop-in for `remfx.models.OpenUnmixModel` (remfx/models.py:259-304)."""

    def __init__(self, n_fft: int = 2048, hop_length: int = 512, n_channels: int = 1, alpha: float = 0.3,
                 sample_rate: int = 22050):
        super().__init__()
        if n_channels != 1:
            raise ValueError("remfx_b200.OpenUnmixModel supports mono audio only (RemFx uses n_channels=1)")
        self.n_channels = n_channels
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.alpha = alpha
        self.register_buffer("window", torch.hann_window(n_fft))
        self.num_bins = n_fft // 2 + 1
        self.sample_rate = sample_rate
        self.model = _OpenUnmixParams(nb_bins=self.num_bins, nb_channels=n_channels)
        self.separator = _SeparatorParams({"other": self.model}, sample_rate, n_fft)
        self._handle: Optional[C.c_void_p] = None
        self._stamp = None
        self._ws: Optional[Tensor] = None

    # ------------------------------------------------------------------ C-ABI handle management
    def _tensors(self):
        core = {k: v for k, v in self.model.state_dict(keep_vars=True).items() if v.dtype == torch.float32}
        core["window"] = self.separator.stft.window
        return core

    def _sync(self, device) -> C.c_void_p:
        tensors = self._tensors()
        stamp = (str(device),) + tuple((k, t.data_ptr(), t._version) for k, t in tensors.items())
        L = _lib.lib()
        if self._handle is not None and stamp == self._stamp:
            return self._handle
        if self._handle is not None:
            # parameters changed: steps still in flight on the library's lane / recurrence / copy streams read the handle's
            # weight buffers, and rfx_umx_load_param overwrites them on the caller's stream -- drain everything first
            for ref in list(self.__dict__.get("_pipes", [])):
                pipe = ref()
                if pipe is not None:
                    pipe.drain()
            for slot in (0, 1):
                _lib.check(L.rfx_umx_wait_host(self._handle, slot), "rfx_umx_wait_host")
            torch.cuda.current_stream(device).synchronize()
        if self._handle is None:
            cfg = _lib.UmxConfig(self.n_fft, self.hop_length, self.model.hidden_size, self.model.nb_layers, 0)
            h = C.c_void_p()
            _lib.check(L.rfx_umx_create(C.byref(cfg), C.byref(h)), "rfx_umx_create")
            self._handle = h
        stream = _lib.cur_stream()
        for k, t in tensors.items():
            if t.device != device:
                raise _lib.RfxError(f"parameter {k} is on {t.device}, input on {device}: call .to(device) first")
            tc = t.detach().contiguous()
            _lib.check(L.rfx_umx_load_param(self._handle, k.encode(), tc.data_ptr(), tc.numel(), stream), f"load {k}")
        _lib.check(L.rfx_umx_finalize(self._handle, stream), "rfx_umx_finalize")
        self._stamp = stamp
        return self._handle

    def __del__(self):
        h = self.__dict__.get("_handle")
        if h is not None:
            self.__dict__["_handle"] = None
            try:
                _lib.lib().rfx_umx_destroy(h)
            except Exception:
                pass

    def set_profiling(self, on: bool, device="cuda:0") -> None:
        """Record cudaEvents between the kernel launches of subsequent sample() calls."""
        with torch.cuda.device(torch.device(device)):
            h = self._sync(torch.device(device))
        _lib.check(_lib.lib().rfx_umx_set_profiling(h, int(on)), "rfx_umx_set_profiling")

    def stage_times_ms(self) -> dict:
        """Per-stage device time (ms) of the last profiled sample(); the stream must be synchronised first."""
        buf = (C.c_float * 32)()
        n = C.c_int()
        _lib.check(_lib.lib().rfx_umx_stage_times(self._handle, buf, 32, C.byref(n)), "rfx_umx_stage_times")
        vals = [buf[i] for i in range(n.value)]
        L = self.model.nb_layers
        names = ["stft", "fc1"] + [f"{k}{l}" for l in range(L) for k in ("wih", "lstm")] + ["fc2", "fc3", "istft"]
        return dict(zip(names, vals))

    def _workspace(self, h, B: int, T: int, device) -> Tensor:
        need = _lib.lib().rfx_umx_workspace_bytes(h, B, T)
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            for slot in (0, 1):  # in-flight host-pipeline calls still use the old workspace
                _lib.check(_lib.lib().rfx_umx_wait_host(h, slot), "rfx_umx_wait_host")
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    # ------------------------------------------------------------------ reference API
    def sample(self, x: Tensor) -> Tensor:
        """(B, 1, T) -> (B, 1, T): `self.separator(x).squeeze(1)` of the reference (models.py:303-304)."""
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"expected input of shape (batch, 1, time), got {tuple(x.shape)}")
        _lib.require_device(x)
        if x.dtype != torch.float32:
            raise ValueError("expected float32 audio")
        x = x.contiguous()
        B, _, T = x.shape
        with torch.cuda.device(x.device):
            h = self._sync(x.device)
            ws = self._workspace(h, B, T, x.device)
            out = torch.empty_like(x)
            rc = _lib.lib().rfx_umx_sample(h, x.data_ptr(), B, T, out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.cur_stream())
            _lib.check(rc, "rfx_umx_sample")
        return out

    def sample_host(self, x_host: Tensor, out_host: Optional[Tensor] = None, device="cuda:0") -> Tensor:
        """End-to-end call on HOST buffers (pinned recommended): H2D, kernels, D2H, stream sync."""
        if x_host.is_cuda or x_host.dim() != 3 or x_host.shape[1] != 1 or x_host.dtype != torch.float32:
            raise ValueError("expected a float32 CPU tensor of shape (batch, 1, time)")
        device = torch.device(device)
        x_host = x_host.contiguous()
        B, _, T = x_host.shape
        if out_host is None:
            out_host = torch.empty_like(x_host, pin_memory=True)
        with torch.cuda.device(device):
            h = self._sync(device)
            ws = self._workspace(h, B, T, device)
            rc = _lib.lib().rfx_umx_sample_host(h, x_host.data_ptr(), B, T, out_host.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _lib.cur_stream())
            _lib.check(rc, "rfx_umx_sample_host")
        return out_host

    def submit_host(self, x_host: Tensor, out_host: Tensor, slot: int = 0, device="cuda:0") -> None:
        """Pipelined `sample_host`: enqueue H2D + kernels + D2H for this batch and return; `wait_host(slot)` blocks until
        `out_host` is complete.  Two slots: submitting slot 1 while slot 0 is in flight overlaps the copies of one batch with
        the kernels of the other.  Both tensors must be pinned and stay alive until the wait."""
        if x_host.is_cuda or x_host.dim() != 3 or x_host.shape[1] != 1 or x_host.dtype != torch.float32 or not x_host.is_contiguous():
            raise ValueError("expected a contiguous float32 CPU tensor of shape (batch, 1, time)")
        if out_host.is_cuda or out_host.shape != x_host.shape or out_host.dtype != torch.float32 or not out_host.is_contiguous():
            raise ValueError("out_host must be a contiguous float32 CPU tensor shaped like x_host")
        if not (x_host.is_pinned() and out_host.is_pinned()):
            raise ValueError("submit_host needs pinned host tensors (pageable memory would make the copies synchronous)")
        device = torch.device(device)
        B, _, T = x_host.shape
        with torch.cuda.device(device):
            h = self._sync(device)
            ws = self._workspace(h, B, T, device)
            rc = _lib.lib().rfx_umx_submit_host(h, int(slot), x_host.data_ptr(), B, T, out_host.data_ptr(), ws.data_ptr(), ws.numel(),
                                                _lib.cur_stream())
            _lib.check(rc, "rfx_umx_submit_host")

    def wait_host(self, slot: int = 0) -> None:
        h = self.__dict__.get("_handle")
        if h:
            _lib.check(_lib.lib().rfx_umx_wait_host(h, int(slot)), "rfx_umx_wait_host")

    def forward(self, batch):
        """(x, target) -> (loss, sep_out) with loss = MRSTFT + 100 * L1 (models.py:294-301)."""
        from .losses import remfx_loss_with_terms

        x, target = batch
        if self.training:
            # training mode (remfx/models.py:294-301 under Lightning's training_step): BatchNorm batch statistics, LSTM dropout,
            # the extra statistics-only pass on spectrogram(x), and a differentiable output
            sep_out = self._forward_train(x)
        else:
            sep_out = self.sample(x)
        loss, self.last_loss_terms = remfx_loss_with_terms(sep_out, target)
        return loss, sep_out

    # ------------------------------------------------------------------ training mode
    def _dropout_masks(self, M: int, device) -> Optional[Tensor]:
        """Inverted-dropout masks of the L - 1 inter-layer dropouts (nn.LSTM(dropout=0.4), umx/openunmix/model.py:62-69), drawn from
        torch's generator; (L - 1, M, hidden) fp32 with entries 0 or 1 / (1 - p).  None when p == 0."""
        p = float(self.model.lstm.dropout)
        L = self.model.nb_layers
        if p <= 0.0 or L < 2:
            return None
        keep = torch.rand(L - 1, M, self.model.hidden_size, device=device) >= p
        return keep.to(torch.float32) / (1.0 - p)

    def _forward_train(self, x: Tensor) -> Tensor:
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"expected input of shape (batch, 1, time), got {tuple(x.shape)}")
        _lib.require_device(x)
        if x.dtype != torch.float32:
            raise ValueError("expected float32 audio")
        x = x.contiguous()
        names = [k for k, v in self.model.named_parameters()]
        params = [v for k, v in self.model.named_parameters()]
        M = x.shape[0] * (x.shape[2] // self.hop_length + 1)
        forced = self.__dict__.get("_forced_masks")  # tests inject (dead-pass masks, real-pass masks)
        masks_dead, masks_real = forced if forced is not None else (self._dropout_masks(M, x.device), self._dropout_masks(M, x.device))
        return _UmxTrainFn.apply(self, names, x, masks_dead, masks_real, *params)

    def _update_running_stats(self, stats: Tensor, M: int) -> None:
        """torch.nn.BatchNorm1d's training-mode bookkeeping from one pass's batch statistics ([mean1, var1, mean2, var2, mean3,
        var3], biased variances): running = (1 - momentum) running + momentum batch, the variance unbiased (M / (M - 1))."""
        o = 0
        unb = float(M) / float(max(M - 1, 1))
        with torch.no_grad():
            for bn in (self.model.bn1, self.model.bn2, self.model.bn3):
                n = bn.num_features
                mean, var = stats[o:o + n], stats[o + n:o + 2 * n]
                o += 2 * n
                if not bn.track_running_stats or bn.running_mean is None:
                    continue
                bn.num_batches_tracked += 1
                mom = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1.0 - mom).add_(mean, alpha=mom)
                bn.running_var.mul_(1.0 - mom).add_(var * unb, alpha=mom)

    def launches_per_call(self) -> int:
        return 5 + 2 * self.model.nb_layers

    def pipeline(self, device="cuda:0") -> "UmxPipeline":
        """Throughput form of `sample` for a stream of equally shaped batches (see UmxPipeline)."""
        return UmxPipeline(self, device)


class _UmxTrainFn(torch.autograd.Function):
    """Open-Unmix training-mode forward + hand-written backward (rfx_umx_forward_train / rfx_umx_backward).  The forward runs the
    reference's two passes: the statistics-only pass on spectrogram(x) (`Y = self.model(X)`, remfx/models.py:296-297 -- its output
    is discarded there too) and the separator pass the loss sees; the BatchNorm running statistics move once per pass, as in the
    reference.  Gradients go to the network's parameters only (the reference detaches the spectrogram it feeds the network)."""

    @staticmethod
    def forward(ctx, owner: "OpenUnmixModel", names, x: Tensor, masks_dead, masks_real, *params: Tensor):
        B, _, T = x.shape
        L = _lib.lib()
        hid, bins = owner.model.hidden_size, owner.num_bins
        with torch.cuda.device(x.device):
            h = owner._sync(x.device)
            need = L.rfx_umx_train_workspace_bytes(h, B, T)
            if need == 0:
                raise ValueError(f"unsupported input size (B={B}, T={T})")
            ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            out = torch.empty_like(x)
            M = B * (T // owner.hop_length + 1)
            for m in (masks_dead, masks_real):
                if m is not None and (m.dtype != torch.float32 or not m.is_contiguous() or m.numel() != (owner.model.nb_layers - 1) * M * hid
                                      or m.device != x.device):
                    raise ValueError("dropout masks must be contiguous float32 CUDA tensors of shape (layers - 1, B * frames, hidden)")
            stats = torch.empty(2, 2 * (hid + hid + bins), dtype=torch.float32, device=x.device)
            # The statistics pass only produces BatchNorm statistics and is independent of the pass the loss sees: it runs on a side
            # stream in its own workspace, beside the separator pass (whose recurrences leave most of the chip idle), and the two
            # join before the running statistics are updated -- in the reference's order, statistics pass first.
            main = torch.cuda.current_stream(x.device)
            side = owner.__dict__.get("_side_stream")
            if side is None or side.device != x.device:
                side = owner.__dict__["_side_stream"] = torch.cuda.Stream(device=x.device)
            rc = L.rfx_umx_train_prepare(h, main.cuda_stream)  # the backward's weight packs (once per parameter change), on `main`
            _lib.check(rc, "rfx_umx_train_prepare")
            ws2 = torch.empty(need, dtype=torch.uint8, device=x.device)
            side.wait_stream(main)  # x, the masks, the parameter upload and the packs are ready
            rc = L.rfx_umx_forward_train(h, x.data_ptr(), B, T, None, ws2.data_ptr(), ws2.numel(),
                                         masks_dead.data_ptr() if masks_dead is not None else None, 1, float(owner.alpha),
                                         stats[0].data_ptr(), side.cuda_stream)
            _lib.check(rc, "rfx_umx_forward_train (statistics pass)")
            rc = L.rfx_umx_forward_train(h, x.data_ptr(), B, T, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                         masks_real.data_ptr() if masks_real is not None else None, 0, float(owner.alpha),
                                         stats[1].data_ptr(), main.cuda_stream)
            _lib.check(rc, "rfx_umx_forward_train")
            main.wait_stream(side)
            for t_ in (ws2, x, stats) + ((masks_dead,) if masks_dead is not None else ()):
                t_.record_stream(side)  # the caching allocator must not hand these blocks out before the side stream is done
        owner._update_running_stats(stats[0], M)
        owner._update_running_stats(stats[1], M)
        ctx.owner, ctx.names = owner, list(names)
        ctx.masks = masks_real  # kept alive: the backward reads them through the pointer the forward recorded
        ctx.save_for_backward(x, ws, *params)
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        x, ws, *params = ctx.saved_tensors
        owner, names = ctx.owner, ctx.names
        B, _, T = x.shape
        L = _lib.lib()
        with torch.cuda.device(x.device):
            # no _sync here: the running statistics moved after the forward, which would re-upload the parameters and drop the
            # tape; the handle still holds the weights the forward used
            h = owner._handle
            from .optim import alloc_param_grads
            grads, _ = alloc_param_grads(params)
            n = len(names)
            keys = (C.c_char_p * n)(*[k.encode() for k in names])
            ptrs = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
            d = dout.detach().to(torch.float32).contiguous()
            rc = L.rfx_umx_backward(h, x.data_ptr(), d.data_ptr(), B, T, keys, ptrs, n, ws.data_ptr(), ws.numel(), _lib.cur_stream())
            _lib.check(rc, "rfx_umx_backward")
        return (None, None, None, None, None, *[g if p.requires_grad else None for g, p in zip(grads, params)])


class UmxPipeline:
    """Multi-lane pipelined `OpenUnmixModel.sample` (rfx_umx_pipe_* in include/remfx_b200.h).

    `push(x, out)` enqueues one batch and returns its sequence number; batch n leaves the pipeline during push(n + depth - 1)
    or `flush()`.  x / out are CUDA tensors, or pinned CPU tensors (the copies then run on the library's copy streams).
    The per-step results are those of `sample` (same kernels in the same order); only the scheduling differs: the LSTM
    recurrences of `depth` consecutive batches run back to back on their own stream while every other kernel of those
    batches runs beside them on the SMs the recurrence does not use.
    """

    def __init__(self, model: OpenUnmixModel, device="cuda:0"):
        import weakref

        self.model = model
        self.device = torch.device(device)
        self._ws: Optional[Tensor] = None
        self._shape = None
        self._keep = {}   # seq -> (x, out): the buffers of every step whose completion has not been OBSERVED yet
        self._outs = {}   # seq -> out for the last _RING steps
        self._last_seq = -1
        with torch.cuda.device(self.device):
            self._h = model._sync(self.device)
        self.depth = _lib.lib().rfx_umx_pipe_depth(self._h)
        model.__dict__.setdefault("_pipes", []).append(weakref.ref(self))

    @staticmethod
    def _check(t: Tensor, name: str) -> bool:
        if t.dim() != 3 or t.shape[1] != 1 or t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError(f"{name}: expected a contiguous float32 tensor of shape (batch, 1, time)")
        if not t.is_cuda and not t.is_pinned():
            raise ValueError(f"{name}: host tensors must be pinned (pageable memory would make the copies synchronous)")
        return not t.is_cuda

    def push(self, x: Tensor, out: Optional[Tensor] = None) -> int:
        x_host = self._check(x, "x")
        if out is None:
            out = torch.empty_like(x, pin_memory=True) if x_host else torch.empty_like(x)
        out_host = self._check(out, "out")
        if out.shape != x.shape:
            raise ValueError("out must be shaped like x")
        for t in (x, out):
            if t.is_cuda and t.device != self.device:
                raise _lib.RfxError(f"tensor on {t.device}, pipeline on {self.device}")
        B, _, T = x.shape
        L = _lib.lib()
        with torch.cuda.device(self.device):
            h = self.model._sync(self.device)
            if self._shape != (B, T):
                self.flush()
                need = L.rfx_umx_pipe_workspace_bytes(h, B, T)
                if self._ws is None or self._ws.numel() < need:
                    torch.cuda.synchronize(self.device)
                    self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
                self._shape = (B, T)
            seq = C.c_longlong(-1)
            rc = L.rfx_umx_pipe_push(h, x.data_ptr(), int(x_host), B, T, out.data_ptr(), int(out_host), self._ws.data_ptr(),
                                     self._ws.numel(), _lib.cur_stream(), C.byref(seq))
            _lib.check(rc, "rfx_umx_pipe_push")
        # The library's lane / recurrence / copy streams are unknown to torch's caching allocators, so the buffers must stay
        # referenced until the step's completion event has actually fired (a push only enqueues: the host can run far ahead).
        self._keep[seq.value] = (x, out)
        self._outs[seq.value] = out   # wait(seq) hands the output back for as long as the library keeps the completion record
        self._outs.pop(seq.value - self._RING, None)
        self._last_seq = seq.value
        self._release_done(seq.value)
        return seq.value

    _RING = 16  # rfx_umx::kRing: completion records are kept for the last 16 steps

    def _release_done(self, newest: int) -> None:
        L = _lib.lib()
        done = C.c_int(0)
        for k in sorted(self._keep):
            if k == newest:
                continue
            if k <= newest - (self._RING - 1):
                # its completion record is about to be recycled: block on it (long finished in any sane schedule)
                _lib.check(L.rfx_umx_pipe_wait(self._h, k), "rfx_umx_pipe_wait")
                del self._keep[k]
                continue
            _lib.check(L.rfx_umx_pipe_query(self._h, k, C.byref(done)), "rfx_umx_pipe_query")
            if done.value:
                del self._keep[k]

    def drain(self) -> None:
        """Flush and block until every step pushed so far is complete (used before the model's weights are re-uploaded)."""
        if self._last_seq < 0:
            return
        self.flush()
        _lib.check(_lib.lib().rfx_umx_pipe_wait(self._h, self._last_seq), "rfx_umx_pipe_wait")
        torch.cuda.current_stream(self.device).synchronize()
        self._keep.clear()   # (_outs stays: wait() on a drained step returns at once)

    def flush(self) -> None:
        """Run the stages still owed to the batches in flight; the current stream then waits for all their outputs."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rfx_umx_pipe_flush(self._h, _lib.cur_stream()), "rfx_umx_pipe_flush")

    def wait(self, seq: int) -> Tensor:
        """Block the host until batch `seq`'s output is complete; returns the output tensor given to push."""
        seq = int(seq)
        if seq not in self._outs:
            raise _lib.RfxError(f"UmxPipeline.wait({seq}): unknown step, or older than the last {self._RING} pushes (completion records "
                                "are kept for that many steps only); keep your own reference to `out` and wait earlier")
        if seq in self._keep:  # completion not observed yet
            _lib.check(_lib.lib().rfx_umx_pipe_wait(self._h, seq), "rfx_umx_pipe_wait")
            del self._keep[seq]
        return self._outs[seq]

    def info(self) -> dict:
        """Schedule facts (valid after the first push): SM partition, recurrence streams, slots per recurrence cluster."""
        a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        _lib.check(_lib.lib().rfx_umx_pipe_info(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), "rfx_umx_pipe_info")
        return {"recurrence_sms": a.value, "other_sms": b.value, "recurrence_streams": c.value, "slots_per_cluster": d.value,
                "partition": "green contexts" if a.value else ("grid caps" if b.value else "none")}

    def set_profiling(self, max_launches: int) -> None:
        """Time the next `max_launches` recurrence launches with cudaEvents on the recurrence stream (0 = off)."""
        _lib.check(_lib.lib().rfx_umx_pipe_set_profiling(self._h, int(max_launches)), "rfx_umx_pipe_set_profiling")

    def recurrence_times_ms(self) -> list:
        """Durations of the profiled recurrence launches (device must be synchronised first)."""
        buf = (C.c_float * 4096)()
        n = C.c_int()
        _lib.check(_lib.lib().rfx_umx_pipe_rec_times(self._h, buf, 4096, C.byref(n)), "rfx_umx_pipe_rec_times")
        return [buf[i] for i in range(n.value)]

    def stream_wait(self, seq: int) -> None:
        """Make the current CUDA stream wait for batch `seq`'s output (no host blocking)."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().rfx_umx_pipe_stream_wait(self._h, int(seq), _lib.cur_stream()), "rfx_umx_pipe_stream_wait")


# ======================================================================================================
# TCN
# ======================================================================================================
class _TCNBlockParams(nn.Module):
    """Parameter layout of remfx/tcn.py:11-46 (TCNBlock.__init__)."""

    def __init__(self, in_ch: int, out_ch: int, kernel_size: int, dilation: int):
        super().__init__()
        self.conv1 = nn.Conv1d(in_ch, out_ch, kernel_size, stride=1, padding=0, dilation=dilation, bias=True)
        self.res = nn.Conv1d(in_ch, out_ch, kernel_size=1, groups=1, stride=1, bias=False)
        self.relu = nn.PReLU(out_ch)


class _TCNParams(nn.Module):
    """Parameter layout of remfx/tcn.py:62-124 (TCN.__init__); same ctor kwargs as cfg/model/tcn.yaml."""

    def __init__(self, ninputs: int = 1, noutputs: int = 1, nblocks: int = 4, channel_growth: int = 0, channel_width: int = 32,
                 kernel_size: int = 13, stack_size: int = 10, dilation_growth: int = 10, condition: bool = False, latent_dim: int = 2,
                 norm_type: str = "identity", causal: bool = False, estimate_loudness: bool = False):
        super().__init__()
        if channel_growth > 1:
            raise ValueError("remfx_b200 TCN supports channel_growth <= 1 (constant channel_width), as in cfg/model/tcn.yaml")
        if condition or estimate_loudness:
            raise ValueError("remfx_b200 TCN does not implement the unused `condition` / `estimate_loudness` options")
        self.ninputs, self.noutputs, self.nblocks = ninputs, noutputs, nblocks
        self.channel_width, self.kernel_size = channel_width, kernel_size
        self.stack_size, self.dilation_growth, self.causal = stack_size, dilation_growth, causal
        self.process_blocks = nn.ModuleList()
        for n in range(nblocks):
            in_ch = channel_width if n > 0 else ninputs
            self.process_blocks.append(_TCNBlockParams(in_ch, channel_width, kernel_size, dilation_growth ** (n % stack_size)))
        self.output = nn.Conv1d(channel_width, noutputs, kernel_size=1)
        self.receptive_field = self.compute_receptive_field()

    def compute_receptive_field(self) -> int:
        """remfx/tcn.py:132-138."""
        rf = self.kernel_size
        for n in range(1, self.nblocks):
            rf += (self.kernel_size - 1) * self.dilation_growth ** (n % self.stack_size)
        return rf


class TCNModel(nn.Module):
    """B200-native drop-in for `remfx.models.TCNModel` (remfx/models.py:370-390)."""

    def __init__(self, sample_rate, num_bins, **kwargs):
        super().__init__()
        self.model = _TCNParams(**kwargs)
        self.sample_rate = sample_rate
        self.num_bins = num_bins
        self._handle: Optional[C.c_void_p] = None
        self._stamp = None
        self._ws: Optional[Tensor] = None

    def _sync(self, device) -> C.c_void_p:
        tensors = dict(self.model.state_dict(keep_vars=True))
        stamp = (str(device),) + tuple((k, t.data_ptr(), t._version) for k, t in tensors.items())
        L = _lib.lib()
        if self._handle is not None and stamp == self._stamp:
            return self._handle
        m = self.model
        if self._handle is None:
            cfg = _lib.TcnConfig(m.ninputs, m.noutputs, m.nblocks, m.channel_width, m.kernel_size, m.stack_size, m.dilation_growth,
                                 int(bool(m.causal)))
            h = C.c_void_p()
            _lib.check(L.rfx_tcn_create(C.byref(cfg), C.byref(h)), "rfx_tcn_create")
            self._handle = h
        stream = _lib.cur_stream()
        for k, t in tensors.items():
            if t.device != device:
                raise _lib.RfxError(f"parameter {k} is on {t.device}, input on {device}: call .to(device) first")
            tc = t.detach().contiguous()
            _lib.check(L.rfx_tcn_load_param(self._handle, k.encode(), tc.data_ptr(), tc.numel(), stream), f"load {k}")
        _lib.check(L.rfx_tcn_finalize(self._handle, stream), "rfx_tcn_finalize")
        self._stamp = stamp
        return self._handle

    def __del__(self):
        h = self.__dict__.get("_handle")
        if h is not None:
            self.__dict__["_handle"] = None
            try:
                _lib.lib().rfx_tcn_destroy(h)
            except Exception:
                pass

    def out_length(self, T: int) -> int:
        return T - (self.model.receptive_field - 1)

    def sample(self, x: Tensor) -> Tensor:
        """(B, 1, T) -> (B, 1, T - receptive_field + 1) (remfx/models.py:388-390)."""
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"expected input of shape (batch, 1, time), got {tuple(x.shape)}")
        _lib.require_device(x)
        if x.dtype != torch.float32:
            raise ValueError("expected float32 audio")
        x = x.contiguous()
        B, _, T = x.shape
        L = _lib.lib()
        with torch.cuda.device(x.device):
            h = self._sync(x.device)
            Lout = L.rfx_tcn_out_length(h, T)
            if Lout <= 0:
                raise ValueError(f"input length {T} is shorter than the receptive field {self.model.receptive_field}")
            need = L.rfx_tcn_workspace_bytes(h, B, T)
            if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            out = torch.empty(B, 1, Lout, dtype=torch.float32, device=x.device)
            rc = L.rfx_tcn_forward(h, x.data_ptr(), B, T, out.data_ptr(), self._ws.data_ptr(), self._ws.numel(), _lib.cur_stream())
            _lib.check(rc, "rfx_tcn_forward")
        return out

    def forward(self, batch):
        """(x, target) -> (loss, output); the target is causal-cropped to the output length (models.py:379-386).

        With autograd enabled and trainable parameters, `output` carries a graph node whose backward is
        `rfx_tcn_backward` (csrc/tcn_bwd.cu), so `loss.backward()` fills every parameter's `.grad` the way the
        reference's Lightning step does (remfx/models.py:217-220)."""
        from .losses import remfx_loss_with_terms
        from .ops import causal_crop

        x, target = batch
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters()):
            output = self._sample_train(x)
        else:
            output = self.sample(x)
        if output.shape[-1] < target.shape[-1]:
            target = causal_crop(target, output.shape[-1])
        loss, self.last_loss_terms = remfx_loss_with_terms(output, target)
        return loss, output

    def _sample_train(self, x: Tensor) -> Tensor:
        if x.dim() != 3 or x.shape[1] != 1:
            raise ValueError(f"expected input of shape (batch, 1, time), got {tuple(x.shape)}")
        _lib.require_device(x)
        if x.dtype != torch.float32:
            raise ValueError("expected float32 audio")
        names = [k for k, _ in self.model.named_parameters()]
        params = [p for _, p in self.model.named_parameters()]
        return _TcnTrainFn.apply(self, names, x.contiguous(), *params)

    def launches_per_call(self) -> int:
        return self.model.nblocks + 1


class _TcnTrainFn(torch.autograd.Function):
    """TCN forward that keeps the block outputs + hand-written backward (rfx_tcn_forward_train / rfx_tcn_backward).

    Gradients are produced for the parameters only: the reference never differentiates with respect to the audio."""

    @staticmethod
    def forward(ctx, owner: "TCNModel", names, x: Tensor, *params: Tensor):
        B, _, T = x.shape
        L = _lib.lib()
        with torch.cuda.device(x.device):
            h = owner._sync(x.device)
            Lout = L.rfx_tcn_out_length(h, T)
            if Lout <= 0:
                raise ValueError(f"input length {T} is shorter than the receptive field {owner.model.receptive_field}")
            ws = torch.empty(L.rfx_tcn_train_workspace_bytes(h, B, T), dtype=torch.uint8, device=x.device)
            out = torch.empty(B, 1, Lout, dtype=torch.float32, device=x.device)
            rc = L.rfx_tcn_forward_train(h, x.data_ptr(), B, T, out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.cur_stream())
            _lib.check(rc, "rfx_tcn_forward_train")
        ctx.owner, ctx.names = owner, list(names)
        ctx.save_for_backward(x, out, ws, *params)
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        x, out, ws, *params = ctx.saved_tensors
        owner, names = ctx.owner, ctx.names
        B, _, T = x.shape
        L = _lib.lib()
        with torch.cuda.device(x.device):
            h = owner._sync(x.device)  # parameters unchanged since the forward: a no-op stamp check
            from .optim import alloc_param_grads
            grads, _ = alloc_param_grads(params)
            n = len(names)
            keys = (C.c_char_p * n)(*[k.encode() for k in names])
            ptrs = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
            d = dout.detach().to(torch.float32).contiguous()
            rc = L.rfx_tcn_backward(h, x.data_ptr(), out.data_ptr(), d.data_ptr(), B, T, keys, ptrs, n, ws.data_ptr(), ws.numel(),
                                    _lib.cur_stream())
            _lib.check(rc, "rfx_tcn_backward")
        return (None, None, None, *[g if p.requires_grad else None for g, p in zip(grads, params)])


# ======================================================================================================
# Hybrid Demucs
# ======================================================================================================
class DemucsModel(nn.Module):
    """B200-native drop-in for `remfx.models.DemucsModel` (remfx/models.py:308-324).

    `self.model` is a `torchaudio.models.HDemucs` instance used ONLY as the parameter container (it gives the
    reference's 397 state_dict keys and its initialisation, incl. `_rescale_module`); its forward is never called --
    the computation is csrc/hdemucs.cu.  kwargs as in cfg/model/demucs.yaml:12-16."""

    def __init__(self, sample_rate, **kwargs) -> None:
        super().__init__()
        from torchaudio.models import HDemucs

        self.model = HDemucs(**kwargs)
        self.num_bins = kwargs["nfft"] // 2 + 1
        self.sample_rate = sample_rate
        self._kw = dict(kwargs)
        self.register_buffer("_hann", torch.hann_window(kwargs["nfft"]), persistent=False)
        self._handle: Optional[C.c_void_p] = None
        self._stamp = None
        self._ws: Optional[Tensor] = None

    def _cfg(self) -> "_lib.HDemucsConfig":
        kw, m = self._kw, self.model
        g = lambda k, d: kw.get(k, d)  # noqa: E731
        return _lib.HDemucsConfig(m.audio_channels, len(m.sources), m.channels, g("growth", 2), m.nfft, m.depth, m.kernel_size, m.stride,
                                  g("time_stride", 2), m.context, g("context_enc", 0), g("norm_starts", 4), g("norm_groups", 4),
                                  g("dconv_depth", 2), g("dconv_comp", 4), g("dconv_attn", 4), g("dconv_lstm", 4),
                                  float(g("freq_emb", 0.2)), float(g("emb_scale", 10)))

    def _sync(self, device) -> C.c_void_p:
        tensors = {k: v for k, v in self.model.state_dict(keep_vars=True).items() if v.dtype == torch.float32}
        tensors["__window__"] = self._hann
        stamp = (str(device),) + tuple((k, t.data_ptr(), t._version) for k, t in tensors.items())
        L = _lib.lib()
        if self._handle is not None and stamp == self._stamp:
            return self._handle
        if self._handle is None:
            cfg = self._cfg()
            h = C.c_void_p()
            _lib.check(L.rfx_hdemucs_create(C.byref(cfg), C.byref(h)), "rfx_hdemucs_create")
            self._handle = h
        stream = _lib.cur_stream()
        for k, t in tensors.items():
            if t.device != device:
                raise _lib.RfxError(f"parameter {k} is on {t.device}, input on {device}: call .to(device) first")
            tc = t.detach().contiguous()
            _lib.check(L.rfx_hdemucs_load_param(self._handle, k.encode(), tc.data_ptr(), tc.numel(), stream), f"load {k}")
        _lib.check(L.rfx_hdemucs_finalize(self._handle, stream), "rfx_hdemucs_finalize")
        self._stamp = stamp
        return self._handle

    def __del__(self):
        h = self.__dict__.get("_handle")
        if h is not None:
            self.__dict__["_handle"] = None
            try:
                _lib.lib().rfx_hdemucs_destroy(h)
            except Exception:
                pass

    def sample(self, x: Tensor, taps: bool = False) -> Tensor:
        """(B, 1, T) -> (B, 1, T): `self.model(x).squeeze(1)` of the reference (models.py:323-324)."""
        if x.ndim != 3:
            raise ValueError(f"Expected 3D tensor with dimensions (batch, channel, frames). Found: {x.shape}")
        if x.shape[1] != self.model.audio_channels:
            raise ValueError("The channel dimension of input Tensor must match `audio_channels` of HDemucs model. "
                             f"Found:{x.shape[1]}.")
        _lib.require_device(x)
        if x.dtype != torch.float32:
            raise ValueError("expected float32 audio")
        x = x.contiguous()
        B, _, T = x.shape
        L = _lib.lib()
        with torch.cuda.device(x.device):
            h = self._sync(x.device)
            _lib.check(L.rfx_hdemucs_set_taps(h, int(taps)), "rfx_hdemucs_set_taps")
            need = L.rfx_hdemucs_workspace_bytes(h, B, T)
            if need == 0:
                raise ValueError(f"unsupported input size (B={B}, T={T}): {L.rfx_last_error().decode() or 'T must be at least nfft samples'}")
            if self._ws is None or self._ws.numel() < need or self._ws.device != x.device:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            out = torch.empty_like(x)
            rc = L.rfx_hdemucs_forward(h, x.data_ptr(), B, T, out.data_ptr(), self._ws.data_ptr(), self._ws.numel(), _lib.cur_stream())
            _lib.check(rc, "rfx_hdemucs_forward")
        return out

    def tap(self, name: str) -> Tensor:
        """Named intermediate activation of the last `sample(..., taps=True)` as fp32 (B, Y, X, C) channel-last."""
        L = _lib.lib()
        dims = (C.c_int * 4)()
        _lib.check(L.rfx_hdemucs_tap(self._handle, name.encode(), None, 0, dims, _lib.cur_stream()), f"tap {name}")
        shape = tuple(dims)
        out = torch.empty(shape, dtype=torch.float32, device=self._ws.device)
        _lib.check(L.rfx_hdemucs_tap(self._handle, name.encode(), out.data_ptr(), out.numel(), dims, _lib.cur_stream()), f"tap {name}")
        return out

    supports_training = True  # rfx_hdemucs_forward_train / rfx_hdemucs_backward (csrc/hdemucs_bwd.cu)

    def forward(self, batch):
        """(x, target) -> (loss, output) (remfx/models.py:317-321).  With autograd enabled and trainable parameters the output
        carries a graph node whose backward is `rfx_hdemucs_backward`, so `loss.backward()` fills every parameter's `.grad`
        the way the reference's Lightning step does (remfx/models.py:217-220).  HDemucs has no BatchNorm / dropout: train and
        eval mode compute the same function."""
        from .losses import remfx_loss_with_terms

        x, target = batch
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.model.parameters()):
            output = self._sample_train(x)
        else:
            output = self.sample(x)
        loss, self.last_loss_terms = remfx_loss_with_terms(output, target)
        return loss, output

    def _check_input(self, x: Tensor) -> Tensor:
        if x.ndim != 3:
            raise ValueError(f"Expected 3D tensor with dimensions (batch, channel, frames). Found: {x.shape}")
        if x.shape[1] != self.model.audio_channels:
            raise ValueError("The channel dimension of input Tensor must match `audio_channels` of HDemucs model. "
                             f"Found:{x.shape[1]}.")
        _lib.require_device(x)
        if x.dtype != torch.float32:
            raise ValueError("expected float32 audio")
        return x.contiguous()

    def _unused_parameters(self):
        te = getattr(self.model, "time_encoder", None)
        if te is None or len(te) == 0:
            return set()
        last = len(te) - 1
        return {f"time_encoder.{last}.norm1.weight", f"time_encoder.{last}.norm1.bias"}

    def _sample_train(self, x: Tensor) -> Tensor:
        x = self._check_input(x)
        named = [(k, p) for k, p in self.model.named_parameters()]
        return _HDemucsTrainFn.apply(self, [k for k, _ in named], x, *[p for _, p in named])

    def grad_tap(self, name: str) -> Tensor:
        """Debug: gradient of a tapped activation after the last backward, fp32 (B, Y, X, C)."""
        L = _lib.lib()
        dims = (C.c_int * 4)()
        _lib.check(L.rfx_hdemucs_grad_tap(self._handle, name.encode(), None, 0, dims, _lib.cur_stream()), f"grad_tap {name}")
        out = torch.empty(tuple(dims), dtype=torch.float32, device=next(self.model.parameters()).device)
        _lib.check(L.rfx_hdemucs_grad_tap(self._handle, name.encode(), out.data_ptr(), out.numel(), dims, _lib.cur_stream()), f"grad_tap {name}")
        return out

    def inject_grad(self, name: str, grad: Optional[Tensor]) -> None:
        """Debug: substitute `grad` (fp32 CUDA, the tap's (B, Y, X, C) layout; keep it alive) for the computed gradient of tap `name`
        in the following backward calls; None removes the substitution."""
        self.__dict__.setdefault("_injected", {})
        if grad is None:
            self._injected.pop(name, None)
            _lib.check(_lib.lib().rfx_hdemucs_inject_grad(self._handle, name.encode(), None), "inject_grad")
        else:
            g = grad.detach().to(torch.float32).contiguous()
            self._injected[name] = g
            _lib.check(_lib.lib().rfx_hdemucs_inject_grad(self._handle, name.encode(), g.data_ptr()), "inject_grad")

    def launches_per_call(self, B: int = 1, T: int = 262144) -> int:
        return _lib.lib().rfx_hdemucs_launches_per_call(self._handle, B, T) if self._handle is not None else 0


class _HDemucsTrainFn(torch.autograd.Function):
    """Hybrid-Demucs forward that keeps its pre-activations + hand-written backward (rfx_hdemucs_forward_train /
    rfx_hdemucs_backward).  Gradients are produced for the parameters only: the reference never differentiates with respect to
    the audio."""

    @staticmethod
    def forward(ctx, owner: "DemucsModel", names, x: Tensor, *params: Tensor):
        B, _, T = x.shape
        L = _lib.lib()
        with torch.cuda.device(x.device):
            h = owner._sync(x.device)
            need = L.rfx_hdemucs_train_workspace_bytes(h, B, T)
            if need == 0:
                raise ValueError(f"unsupported input size (B={B}, T={T}): T must be a multiple of 1024 ({_lib.lib().rfx_last_error().decode()})")
            ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            out = torch.empty_like(x)
            rc = L.rfx_hdemucs_forward_train(h, x.data_ptr(), B, T, out.data_ptr(), ws.data_ptr(), ws.numel(), _lib.cur_stream())
            _lib.check(rc, "rfx_hdemucs_forward_train")
        ctx.owner, ctx.names = owner, list(names)
        ctx.save_for_backward(x, ws, *params)
        return out

    @staticmethod
    def backward(ctx, dout: Tensor):
        x, ws, *params = ctx.saved_tensors
        owner, names = ctx.owner, ctx.names
        B, _, T = x.shape
        L = _lib.lib()
        with torch.cuda.device(x.device):
            h = owner._sync(x.device)  # parameters unchanged since the forward: a no-op stamp check
            from .optim import alloc_param_grads
            grads, zeroed = alloc_param_grads(params)   # views of one zeroed flat bucket when a FusedAdamW owns the parameters
            _lib.check(L.rfx_hdemucs_set_grads_prezeroed(h, 1 if zeroed else 0), "rfx_hdemucs_set_grads_prezeroed")
            n = len(names)
            keys = (C.c_char_p * n)(*[k.encode() for k in names])
            ptrs = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
            d = dout.detach().to(torch.float32).contiguous()
            rc = L.rfx_hdemucs_backward(h, x.data_ptr(), d.data_ptr(), B, T, keys, ptrs, n, ws.data_ptr(), ws.numel(), _lib.cur_stream())
            _lib.check(rc, "rfx_hdemucs_backward")
        # parameters the network never uses get no gradient, like under torch autograd (the "empty" innermost time encoder
        # has a norm1 that its forward skips, TA:_hdemucs.py:159-160), so that the optimiser leaves them alone
        unused = owner._unused_parameters()
        return (None, None, None, *[g if (p.requires_grad and k not in unused) else None for k, g, p in zip(names, grads, params)])
