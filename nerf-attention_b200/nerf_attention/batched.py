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
    real torch scheduler runs on a dummy parameter so the float64 recursion is torch's own."""
    dummy = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([dummy], lr=lr)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=epochs, eta_min=lr * 0.01)
    table = np.empty(epochs, dtype=np.float64)
    for e in range(epochs):
        table[e] = opt.param_groups[0]['lr']
        opt.step()
        sched.step()
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


def fit_many(jobs: list[FitJob], epochs: int = 5000, lr: float = 1e-4, device: str = 'cuda',
             log_every: int = 500, verbose: bool = True, precision: str | None = None,
             betas: tuple[float, float] = (0.9, 0.999), eps: float = 1e-8) -> list[FitResult]:
    """Train all jobs for ``epochs`` full-batch Adam steps; one FitResult per job, in order."""
    global last_stats
    dev = _native.require_cuda(device)
    lib = _native.lib()
    prec = _native.precision_code(precision)
    if not jobs:
        return []
    stats = TransferStats()
    t_wall = time.perf_counter()

    # ---- models: built on the CPU in job order so a seeded run matches the reference's stream
    for job in jobs:
        if job.kv_tensor.dim() != 2:
            raise ValueError(f'kv_tensor must be (seq_len, d_head), got {tuple(job.kv_tensor.shape)}')
        if job.model is None:
            job.model = SIREN(job.config, out_features=job.kv_tensor.shape[1])
    n_params = [job.model.count_parameters() for job in jobs]

    with torch.cuda.device(dev):
        # ---- inputs: each distinct tensor / position vector goes up once
        uploaded: dict[tuple, torch.Tensor] = {}
        targets, positions = [], []
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
            targets.append(uploaded[key])
            n = t.shape[0]
            if n not in pos_cache:
                pos_cache[n] = torch.linspace(0, 1, n).to(dev)      # CPU linspace, siren.py:82
                stats.h2d_bytes += 4 * n
            positions.append(pos_cache[n])

        # ---- packed weights: one pinned staging buffer, one copy
        params = _Packer(n_params, dev)
        staging = torch.empty(params.total, dtype=torch.float32, pin_memory=True)
        for i, job in enumerate(jobs):
            pack_model(job.model, staging[params.offsets[i]: params.offsets[i] + n_params[i]])
        params.buf.copy_(staging, non_blocking=True)
        stats.h2d_bytes += params.total * 4
        adam_m = _Packer(n_params, dev, zero=True)
        adam_v = _Packer(n_params, dev, zero=True)

        seq = [job.kv_tensor.shape[0] for job in jobs]
        dh = [job.kv_tensor.shape[1] for job in jobs]
        losses = _Packer([epochs] * len(jobs), dev)
        cos = _Packer(seq, dev)
        ppm = _Packer(seq, dev)
        scal = _Packer([8] * len(jobs), dev)
        mean = _Packer(dh, dev)
        std = _Packer(dh, dev)

        fits = (_native.NaFit * len(jobs))()
        for i, job in enumerate(jobs):
            f = fits[i]
            f.N, f.D = seq[i], dh[i]
            f.H, f.L = job.config.hidden_features, job.config.hidden_layers
            f.omega0, f.flags = job.config.omega_0, 0
            f.positions, f.targets = positions[i].data_ptr(), targets[i].data_ptr()
            f.mean, f.std = mean.ptr(i), std.ptr(i)
            f.params, f.adam_m, f.adam_v = params.ptr(i), adam_m.ptr(i), adam_v.ptr(i)
            f.losses, f.cos_sims, f.per_pos_mse, f.scalars = losses.ptr(i), cos.ptr(i), ppm.ptr(i), scal.ptr(i)

        need = ctypes.c_size_t(0)
        _native.check(lib.nerfattn_fit_workspace_bytes(fits, len(jobs), prec, ctypes.byref(need)),
                      'nerfattn_fit_workspace_bytes')
        workspace = torch.empty(need.value, dtype=torch.uint8, device=dev)
        table = lr_schedule(epochs, lr)
        stats.setup_seconds = time.perf_counter() - t_wall

        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        _native.check(lib.nerfattn_fit_batched(
            fits, len(jobs), epochs, table.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
            betas[0], betas[1], eps, 0, prec, workspace.data_ptr(), need.value,
            _native.stream_handle()), 'nerfattn_fit_batched')
        ev1.record()

        # ---- results back (pinned, async, one sync)
        def fetch(p: _Packer) -> torch.Tensor:
            host = torch.empty(p.total, dtype=torch.float32, pin_memory=True)
            host.copy_(p.buf, non_blocking=True)
            stats.d2h_bytes += p.total * 4
            return host
        h_losses, h_cos, h_ppm, h_scal, h_mean, h_std = (fetch(p) for p in (losses, cos, ppm, scal, mean, std))
        torch.cuda.current_stream().synchronize()
        stats.gpu_seconds = ev0.elapsed_time(ev1) / 1e3
        del workspace

    # ---- unpack
    flops = [job.config.flops_per_epoch(seq[i], dh[i]) for i, job in enumerate(jobs)]
    total_flops = float(sum(flops)) or 1.0
    results: list[FitResult] = []
    for i, job in enumerate(jobs):
        adopt_packed(job.model, params.view(i, n_params[i]))
        job.model.eval()
        sc = h_scal[scal.offsets[i]: scal.offsets[i] + 8]
        n, d = seq[i], dh[i]
        fit_losses = h_losses[losses.offsets[i]: losses.offsets[i] + epochs].tolist()
        raw = n * d * 2                                   # fp16 KV baseline, siren.py:127
        size = job.model.size_bytes()
        if verbose and epochs:
            step = max(int(log_every), 1)
            for e in range(step, epochs + 1, step):
                print(f"  Epoch {e}/{epochs} | NormMSE: {fit_losses[e - 1]:.6f}")
        results.append(FitResult(
            model=job.model, config=job.config,
            target_mean=h_mean[mean.offsets[i]: mean.offsets[i] + d].clone().unsqueeze(0),
            target_std=h_std[std.offsets[i]: std.offsets[i] + d].clone().unsqueeze(0),
            losses=fit_losses,
            final_mse=float(sc[0]), final_cosine_mean=float(sc[1]),
            final_cosine_min=float(sc[2]), final_cosine_std=float(sc[3]),
            per_pos_mse=h_ppm[ppm.offsets[i]: ppm.offsets[i] + n].numpy().copy(),
            cosine_sims=h_cos[cos.offsets[i]: cos.offsets[i] + n].numpy().copy(),
            compression_ratio=raw / size, raw_size_bytes=raw, siren_size_bytes=size,
            # the sweep trains concurrently: a fit's time is its FLOP share of the batch
            train_time_seconds=stats.gpu_seconds * flops[i] / total_flops,
            seq_len=n, d_head=d, num_parameters=n_params[i],
        ))
    stats.wall_seconds = time.perf_counter() - t_wall
    last_stats = stats
    return results
