"""Batched SIREN fitting: every job of a sweep trained at once on the GPU.

Replaces the serial ``for layer: for head: for kv: for config: fit_siren(...)``
loop nest of the reference (nerf_attention/fit.py:54-76).  Host side only packs
and unpacks; all arithmetic runs in libnerfattn.so (include/nerfattn.h).

Host/device traffic of one call (counted in ``TransferStats``):
  H2D  every distinct KV tensor once, positions once per distinct seq_len, the
       packed initial weights of all jobs in one copy
  D2H  losses [jobs, epochs], per-position CosSim / MSE, the 4 scalars, mean/std
       (trained weights stay on the device: the returned models own views of them)
"""

from __future__ import annotations

import ctypes
import functools
import time
from dataclasses import dataclass, field

import numpy as np
import torch

from nerf_attention import _native
from nerf_attention.siren import SIREN
from nerf_attention.types import FitResult, SIRENConfig

_ALIGN = 64   # floats; keeps every packed vector 256-byte aligned for 128-bit loads


@dataclass
class FitJob:
    kv_tensor: torch.Tensor                 # [seq_len, d_head] fp32, CPU or CUDA
    config: SIRENConfig
    model: SIREN | None = None              # pre-built (seeded) model; built in job order if None
    name: str = ''
    # Optional: kv_tensor is ALREADY (t - mean) / std and these are the statistics (NA_FIT_TARGETS_PRENORMALISED):
    # the library skips its normalisation pass and reports the final metrics against kv_tensor * std + mean.
    target_mean: torch.Tensor | None = None  # [d_head] or [1, d_head]
    target_std: torch.Tensor | None = None


@dataclass
class TransferStats:
    h2d_bytes: int = 0
    d2h_bytes: int = 0
    gpu_seconds: float = 0.0               # CUDA-event time of the native call
    wall_seconds: float = 0.0
    setup_seconds: float = 0.0             # host: model construction + packing


last_stats = TransferStats()


def lr_schedule(epochs: int, lr: float) -> np.ndarray:
    """lr seen by optimizer.step() at each epoch: Adam(lr) + CosineAnnealingLR(T_max=epochs,
    eta_min=0.01*lr), stepped after the optimizer (reference siren.py:90-93,103-104).  The
    real torch scheduler runs on a dummy parameter so the float64 recursion is torch's own.
    Stepping it takes ~60 ms for 2000 epochs (most of a batched call's host set-up), so the
    table is computed once per (epochs, lr); callers get a read-only array."""
    return _lr_schedule_cached(int(epochs), float(lr))


@functools.lru_cache(maxsize=32)
def _lr_schedule_cached(epochs: int, lr: float) -> np.ndarray:
    dummy = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([dummy], lr=lr)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=epochs, eta_min=lr * 0.01)
    table = np.empty(epochs, dtype=np.float64)
    for e in range(epochs):
        table[e] = opt.param_groups[0]['lr']
        opt.step()
        sched.step()
    table.setflags(write=False)
    return table


def _round_up(n: int, a: int = _ALIGN) -> int:
    return (n + a - 1) // a * a


class _Packer:
    """Flat fp32 device buffers with one aligned slot per job."""

    def __init__(self, sizes: list[int], device: torch.device, zero: bool = False):
        self.offsets = []
        total = 0
        for s in sizes:
            self.offsets.append(total)
            total += _round_up(s)
        self.total = max(total, _ALIGN)
        self.buf = (torch.zeros if zero else torch.empty)(self.total, dtype=torch.float32, device=device)

    def ptr(self, i: int) -> int:
        return self.buf.data_ptr() + 4 * self.offsets[i]

    def view(self, i: int, n: int) -> torch.Tensor:
        return self.buf[self.offsets[i]: self.offsets[i] + n]


def pack_model(model: SIREN, out: torch.Tensor) -> None:
    """Write state_dict-order weights into a flat host vector (nerfattn.h params layout)."""
    off = 0
    for p in model.packed_parameters():
        n = p.numel()
        out[off:off + n].copy_(p.detach().reshape(-1))
        off += n


def adopt_packed(model: SIREN, flat: torch.Tensor) -> None:
    """Make the model's parameters views of ``flat`` (device) -- zero-copy 'model.to(device)'."""
    off = 0
    for p in model.packed_parameters():
        n = p.numel()
        p.data = flat[off:off + n].view(p.shape)
        off += n


class FitBatch:
    """One batched fit, split into its phases so that callers (bench.py) can time them apart:
    ``__init__`` uploads and packs (H2D), ``launch`` enqueues the native call, ``reset`` restores
    the initial weights / zero Adam moments on the device, ``collect`` reads results back (D2H)
    and builds the FitResults.  ``fit_many`` is ``FitBatch(...).launch().collect()``."""

    def __init__(self, jobs: list[FitJob], epochs: int = 5000, lr: float = 1e-4, device: str = 'cuda',
                 precision: str | None = None, betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 keep_initial: bool = False, progress_every: int = 0, lib=None):
        self.dev = dev = _native.require_cuda(device)
        self.lib = lib if lib is not None else _native.lib()      # `lib`: bench.py's profiling build
        self.prec = _native.precision_code(precision)
        self.jobs, self.epochs, self.betas, self.eps = jobs, epochs, betas, eps
        self.stats = stats = TransferStats()
        t_wall = time.perf_counter()

        # ---- models: built on the CPU in job order so a seeded run matches the reference's stream
        for job in jobs:
            if job.kv_tensor.dim() != 2:
                raise ValueError(f'kv_tensor must be (seq_len, d_head), got {tuple(job.kv_tensor.shape)}')
            if job.model is None:
                job.model = SIREN(job.config, out_features=job.kv_tensor.shape[1])
        self.n_params = n_params = [job.model.count_parameters() for job in jobs]
        self.seq = seq = [job.kv_tensor.shape[0] for job in jobs]
        self.dh = dh = [job.kv_tensor.shape[1] for job in jobs]

        with torch.cuda.device(dev):
            # ---- inputs: each distinct tensor / position vector goes up once
            uploaded: dict[tuple, torch.Tensor] = {}
            self.targets, self.positions = [], []
            pos_cache: dict[int, torch.Tensor] = {}
            for job in jobs:
                t = job.kv_tensor
                key = (t.data_ptr(), tuple(t.shape), tuple(t.stride()), str(t.device))
                if key not in uploaded:
                    src = t.detach()
                    if src.dtype != torch.float32:
                        src = src.float()
                    if src.device != dev:
                        stats.h2d_bytes += src.numel() * 4
                    uploaded[key] = src.to(dev, non_blocking=True).contiguous()
                self.targets.append(uploaded[key])
                n = t.shape[0]
                if n not in pos_cache:
                    pos_cache[n] = torch.linspace(0, 1, n).to(dev)      # CPU linspace, siren.py:82
                    stats.h2d_bytes += 4 * n
                self.positions.append(pos_cache[n])

            # ---- packed weights: one pinned staging buffer, one copy
            self.params = params = _Packer(n_params, dev)
            staging = torch.empty(params.total, dtype=torch.float32, pin_memory=True)
            for i, job in enumerate(jobs):
                pack_model(job.model, staging[params.offsets[i]: params.offsets[i] + n_params[i]])
            params.buf.copy_(staging, non_blocking=True)
            stats.h2d_bytes += params.total * 4
            self.initial = params.buf.clone() if keep_initial else None
            self.adam_m = _Packer(n_params, dev, zero=True)
            self.adam_v = _Packer(n_params, dev, zero=True)
            self.losses = _Packer([epochs] * len(jobs), dev)
            self.cos = _Packer(seq, dev)
            self.ppm = _Packer(seq, dev)
            self.scal = _Packer([8] * len(jobs), dev)
            self.mean = _Packer(dh, dev)
            self.std = _Packer(dh, dev)
            for i, job in enumerate(jobs):                     # pre-normalised targets bring their statistics along
                if (job.target_mean is None) != (job.target_std is None):
                    raise ValueError('target_mean and target_std must be given together')
                if job.target_mean is not None:
                    self.mean.view(i, dh[i]).copy_(job.target_mean.reshape(-1).float(), non_blocking=True)
                    self.std.view(i, dh[i]).copy_(job.target_std.reshape(-1).float(), non_blocking=True)
                    stats.h2d_bytes += 8 * dh[i]

            self.fits = fits = (_native.NaFit * len(jobs))()
            for i, job in enumerate(jobs):
                f = fits[i]
                f.N, f.D = seq[i], dh[i]
                f.H, f.L = job.config.hidden_features, job.config.hidden_layers
                f.omega0, f.flags = job.config.omega_0, (_native.FIT_TARGETS_PRENORMALISED if job.target_mean is not None else 0)
                f.positions, f.targets = self.positions[i].data_ptr(), self.targets[i].data_ptr()
                f.mean, f.std = self.mean.ptr(i), self.std.ptr(i)
                f.params, f.adam_m, f.adam_v = params.ptr(i), self.adam_m.ptr(i), self.adam_v.ptr(i)
                f.losses, f.cos_sims = self.losses.ptr(i), self.cos.ptr(i)
                f.per_pos_mse, f.scalars = self.ppm.ptr(i), self.scal.ptr(i)

            need = ctypes.c_size_t(0)
            _native.check(self.lib.nerfattn_fit_workspace_bytes(fits, len(jobs), self.prec, ctypes.byref(need)),
                          'nerfattn_fit_workspace_bytes')
            self.workspace_bytes = need.value
            self.workspace = torch.empty(need.value, dtype=torch.uint8, device=dev)
            self.table = lr_schedule(epochs, lr)
            # progress metrics (siren.py:107-115): [epochs // every][fit][RealMSE, CosSim]
            self.progress_every = int(progress_every) if progress_every and epochs >= progress_every > 0 else 0
            self.progress = (torch.zeros((epochs // self.progress_every) * len(jobs) * 2, dtype=torch.float32, device=dev)
                             if self.progress_every else None)
        self.flops = [job.config.flops_per_epoch(seq[i], dh[i]) for i, job in enumerate(jobs)]
        stats.setup_seconds = time.perf_counter() - t_wall
        self._t_wall = t_wall

    def launch(self) -> 'FitBatch':
        """Enqueue the whole fit on the current stream (asynchronous)."""
        with torch.cuda.device(self.dev):
            _native.check(self.lib.nerfattn_fit_batched_ex(
                self.fits, len(self.jobs), self.epochs,
                self.table.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                self.betas[0], self.betas[1], self.eps, 0, self.prec, self.progress_every,
                self.progress.data_ptr() if self.progress is not None else None, self.workspace.data_ptr(),
                self.workspace_bytes, _native.stream_handle()), 'nerfattn_fit_batched_ex')
        return self

    def reset(self) -> None:
        """Device-side: initial weights back, Adam moments to zero (needs keep_initial=True)."""
        self.params.buf.copy_(self.initial, non_blocking=True)
        self.adam_m.buf.zero_()
        self.adam_v.buf.zero_()

    def launches_per_call(self) -> int:
        return int(self.lib.nerfattn_fit_launch_count(self.fits, len(self.jobs), self.epochs, self.prec))

    def collect(self, verbose: bool = False, log_every: int = 500, keep_optimizer_state: bool = False,
                gpu_seconds: float | None = None) -> list[FitResult]:
        """D2H of losses / metrics (pinned, async, one sync) and the FitResult records."""
        stats, jobs, epochs = self.stats, self.jobs, self.epochs

        def fetch(p: _Packer) -> torch.Tensor:
            host = torch.empty(p.total, dtype=torch.float32, pin_memory=True)
            host.copy_(p.buf, non_blocking=True)
            stats.d2h_bytes += p.total * 4
            return host
        with torch.cuda.device(self.dev):
            h_losses, h_cos, h_ppm, h_scal, h_mean, h_std = (
                fetch(p) for p in (self.losses, self.cos, self.ppm, self.scal, self.mean, self.std))
            h_prog = None
            if self.progress is not None:
                h_prog = torch.empty(self.progress.numel(), dtype=torch.float32, pin_memory=True)
                h_prog.copy_(self.progress, non_blocking=True)
                stats.d2h_bytes += self.progress.numel() * 4
            torch.cuda.current_stream().synchronize()
        if h_prog is not None:
            h_prog = h_prog.view(-1, len(jobs), 2)
        if gpu_seconds is not None:
            stats.gpu_seconds = gpu_seconds
        total_flops = float(sum(self.flops)) or 1.0
        results: list[FitResult] = []
        for i, job in enumerate(jobs):
            n, d, p = self.seq[i], self.dh[i], self.n_params[i]
            adopt_packed(job.model, self.params.view(i, p))
            job.model.eval()
            if keep_optimizer_state:       # flat Adam moments in params order (tests, warm restarts)
                job.model.adam_state = (self.adam_m.view(i, p), self.adam_v.view(i, p))
            sc = h_scal[self.scal.offsets[i]: self.scal.offsets[i] + 8]
            fit_losses = h_losses[self.losses.offsets[i]: self.losses.offsets[i] + epochs].tolist()
            raw = n * d * 2                                   # fp16 KV baseline, siren.py:127
            size = job.model.size_bytes()
            progress = []
            if h_prog is not None:                           # the reference's progress line, siren.py:112-115
                step = self.progress_every
                for k, e in enumerate(range(step, epochs + 1, step)):
                    progress.append((e, fit_losses[e - 1], float(h_prog[k, i, 0]), float(h_prog[k, i, 1])))
                    if verbose:
                        print(f"  Epoch {e}/{epochs} | NormMSE: {fit_losses[e - 1]:.6f} | "
                              f"RealMSE: {progress[-1][2]:.6f} | CosSim: {progress[-1][3]:.4f}")
            elif verbose and epochs:
                step = max(int(log_every), 1)
                for e in range(step, epochs + 1, step):
                    print(f"  Epoch {e}/{epochs} | NormMSE: {fit_losses[e - 1]:.6f}")
            results.append(FitResult(
                model=job.model, config=job.config,
                target_mean=h_mean[self.mean.offsets[i]: self.mean.offsets[i] + d].clone().unsqueeze(0),
                target_std=h_std[self.std.offsets[i]: self.std.offsets[i] + d].clone().unsqueeze(0),
                losses=fit_losses,
                final_mse=float(sc[0]), final_cosine_mean=float(sc[1]),
                final_cosine_min=float(sc[2]), final_cosine_std=float(sc[3]),
                per_pos_mse=h_ppm[self.ppm.offsets[i]: self.ppm.offsets[i] + n].numpy().copy(),
                cosine_sims=h_cos[self.cos.offsets[i]: self.cos.offsets[i] + n].numpy().copy(),
                compression_ratio=raw / size, raw_size_bytes=raw, siren_size_bytes=size,
                # the sweep trains concurrently: a fit's time is its FLOP share of the batch
                train_time_seconds=stats.gpu_seconds * self.flops[i] / total_flops,
                seq_len=n, d_head=d, num_parameters=p,
            ))
            results[-1].progress = progress      # [(epoch, NormMSE, RealMSE, CosSim)], siren.py:107-115
        stats.wall_seconds = time.perf_counter() - self._t_wall
        return results


def fit_many(jobs: list[FitJob], epochs: int = 5000, lr: float = 1e-4, device: str = 'cuda',
             log_every: int = 500, verbose: bool = True, precision: str | None = None,
             betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
             keep_optimizer_state: bool = False, progress: bool | None = None) -> list[FitResult]:
    """Train all jobs for ``epochs`` full-batch Adam steps; one FitResult per job, in order."""
    global last_stats
    if not jobs:
        _native.require_cuda(device)
        return []
    batch = FitBatch(jobs, epochs=epochs, lr=lr, device=device, precision=precision, betas=betas, eps=eps,
                     progress_every=log_every if (verbose if progress is None else progress) else 0)
    with torch.cuda.device(batch.dev):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        batch.launch()
        ev1.record()
        ev1.synchronize()
        gpu_seconds = ev0.elapsed_time(ev1) / 1e3
    results = batch.collect(verbose=verbose, log_every=log_every, keep_optimizer_state=keep_optimizer_state,
                            gpu_seconds=gpu_seconds)
    last_stats = batch.stats
    return results
