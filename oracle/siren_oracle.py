"""CPU oracle for the SIREN fit / reconstruction hot path.  TEST INFRASTRUCTURE ONLY.

This module restates, in plain PyTorch on the CPU, what the reference computes
for the path BASELINE.json names.  It exists so that the CUDA kernels can be
checked against something, and so that ``bench.py`` can time "the reference's
CPU path" on a box where /root/reference is not mounted.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline / --impl reference)
may import it.  The product package never does: it fails loudly without the
CUDA library.

Parity status: PINNED.  The reference ships no golden vectors or tests
(SURVEY.md 4), and its arithmetic lives in torch (unpinned ``torch>=2.1.0``,
reference pyproject.toml:7; the version that defines "reference results" here
is torch 2.11.0).  The oracle is therefore pinned against outputs of the real
reference imported in the build container: ``tests/golden/make_golden.py``
runs reference ``fit_siren``/``SIREN`` under fixed seeds and commits the
results under ``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.

Each function cites the reference lines it follows (paths relative to
/root/reference).
"""

from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class OracleFit:
    """Everything reference ``fit_siren`` returns, minus the nn.Module wrapper."""

    state: dict[str, torch.Tensor]          # state_dict-keyed weights after training
    target_mean: torch.Tensor               # [1, D]
    target_std: torch.Tensor                # [1, D]
    losses: list[float]
    final_mse: float
    final_cosine_mean: float
    final_cosine_min: float
    final_cosine_std: float
    per_pos_mse: np.ndarray
    cosine_sims: np.ndarray
    compression_ratio: float
    raw_size_bytes: int
    siren_size_bytes: int
    num_parameters: int
    train_time_seconds: float
    progress: list[tuple[int, float, float, float]] = field(default_factory=list)


# --------------------------------------------------------------------------
# model: init + forward
# --------------------------------------------------------------------------

def _fresh_linear(fan_in: int, fan_out: int, bound: float) -> tuple[torch.Tensor, torch.Tensor]:
    """One ``nn.Linear`` followed by the SIREN re-draw.

    nerf_attention/siren.py:23-31 (sine layers) and :51-55 (output layer):
    ``nn.Linear(in, out)`` first consumes the generator for its default
    kaiming/bias init, then weight and bias are both overwritten with
    U(-bound, bound) -- weight first, bias second.  Going through a real
    nn.Linear keeps the CPU generator stream identical to the reference's.
    """
    lin = torch.nn.Linear(fan_in, fan_out)
    with torch.no_grad():
        lin.weight.uniform_(-bound, bound)
        lin.bias.uniform_(-bound, bound)
    return lin.weight.detach().clone(), lin.bias.detach().clone()


def init_state(hidden_features: int, hidden_layers: int, omega_0: float,
               out_features: int) -> dict[str, torch.Tensor]:
    """Seed-for-seed equivalent of ``SIREN(config, out_features).state_dict()``.

    Layer order and bounds: nerf_attention/siren.py:43-58.  First layer bound
    is 1/in_features (=1), every other layer (output layer included) uses
    sqrt(6/in)/omega_0.
    """
    h = hidden_features
    state: dict[str, torch.Tensor] = {}
    w, b = _fresh_linear(1, h, 1.0 / 1)
    state['network.0.linear.weight'], state['network.0.linear.bias'] = w, b
    hb = math.sqrt(6.0 / h) / omega_0
    for i in range(1, hidden_layers + 1):
        w, b = _fresh_linear(h, h, hb)
        state[f'network.{i}.linear.weight'], state[f'network.{i}.linear.bias'] = w, b
    w, b = _fresh_linear(h, out_features, hb)
    k = hidden_layers + 1
    state[f'network.{k}.weight'], state[f'network.{k}.bias'] = w, b
    return state


def _layers(state: dict[str, torch.Tensor]) -> tuple[list[tuple[torch.Tensor, torch.Tensor]],
                                                      tuple[torch.Tensor, torch.Tensor]]:
    sine = []
    i = 0
    while f'network.{i}.linear.weight' in state:
        sine.append((state[f'network.{i}.linear.weight'], state[f'network.{i}.linear.bias']))
        i += 1
    return sine, (state[f'network.{i}.weight'], state[f'network.{i}.bias'])


def forward(state: dict[str, torch.Tensor], omega_0: float, x: torch.Tensor) -> torch.Tensor:
    """``SIREN.forward``: sin(omega_0 * Linear(x)) per sine layer, then a plain Linear.

    nerf_attention/siren.py:33-34 and :60-61.
    """
    sine, (wf, bf) = _layers(state)
    h = x
    for w, b in sine:
        h = torch.sin(omega_0 * F.linear(h, w, b))
    return F.linear(h, wf, bf)


def forward_exact(state: dict[str, torch.Tensor], omega_0: float, x: torch.Tensor) -> torch.Tensor:
    """The same fp32 weights and fp32 inputs evaluated in float64: the function that every fp32 evaluation
    (the reference's torch CPU arithmetic, the CUDA kernels) approximates.  Parity tests use it to tell the
    kernel's rounding error from the reference's own (torch's CPU fp32 path is usually within 1e-6 of it, but
    has been seen 3e-5 off in the first evaluation of a process that had just initialised CUDA)."""
    return forward({k: v.double() for k, v in state.items()}, omega_0, x.double())


def count_parameters(state: dict[str, torch.Tensor]) -> int:
    """nerf_attention/siren.py:63-64."""
    return sum(int(t.numel()) for t in state.values())


# --------------------------------------------------------------------------
# training
# --------------------------------------------------------------------------

def positions_for(seq_len: int) -> torch.Tensor:
    """nerf_attention/siren.py:82 -- CPU linspace, *not* i/(N-1) in fp32."""
    return torch.linspace(0, 1, seq_len).unsqueeze(1)


def normalise(targets: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Per-dimension standardisation, nerf_attention/siren.py:85-87.

    std is the unbiased estimator, clamped from below at 1e-3.
    """
    mean = targets.mean(dim=0, keepdim=True)
    std = targets.std(dim=0, keepdim=True).clamp(min=1e-3)
    return mean, std, (targets - mean) / std


def lr_table(epochs: int, lr: float) -> list[float]:
    """Learning rate used by optimizer.step() of each epoch.

    nerf_attention/siren.py:90-93,103-104: Adam(lr) + CosineAnnealingLR(T_max=epochs,
    eta_min=0.01*lr), scheduler stepped after the optimizer, so epoch 0 runs at
    ``lr``.  The real scheduler (recursive float64 form) is instantiated on a
    dummy parameter so the table is whatever torch produces.
    """
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=lr)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=epochs, eta_min=lr * 0.01)
    out = []
    for _ in range(epochs):
        out.append(float(opt.param_groups[0]['lr']))
        opt.step()
        sched.step()
    return out


def final_metrics(pred_norm: torch.Tensor, targets: torch.Tensor, mean: torch.Tensor,
                  std: torch.Tensor) -> dict:
    """De-normalised quality metrics, nerf_attention/siren.py:119-125,137-139."""
    pred_real = pred_norm * std + mean
    cos = F.cosine_similarity(pred_real, targets, dim=1)
    return {
        'final_mse': F.mse_loss(pred_real, targets).item(),
        'cosine_sims': cos,
        'per_pos_mse': ((pred_real - targets) ** 2).mean(dim=1),
        'final_cosine_mean': cos.mean().item(),
        'final_cosine_min': cos.min().item(),
        'final_cosine_std': cos.std().item(),
    }


def fit(kv_tensor: torch.Tensor, hidden_features: int, hidden_layers: int, omega_0: float,
        epochs: int = 5000, lr: float = 1e-4, device: str = 'cpu', log_every: int = 500,
        init: dict[str, torch.Tensor] | None = None) -> OracleFit:
    """Restatement of reference ``fit_siren`` (nerf_attention/siren.py:70-149).

    ``init`` lets a test hand in the exact initial weights given to the CUDA
    path; when None the model is drawn from the global CPU generator exactly
    as the reference does (model built on CPU, then moved, :89).
    """
    seq_len, d_head = kv_tensor.shape
    x = positions_for(seq_len).to(device)
    targets = kv_tensor.to(device)
    mean, std, t_norm = normalise(targets)

    state0 = init if init is not None else init_state(hidden_features, hidden_layers, omega_0, d_head)
    params = {k: v.detach().clone().to(device).requires_grad_(True) for k, v in state0.items()}
    opt = torch.optim.Adam(list(params.values()), lr=lr)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=epochs, eta_min=lr * 0.01)

    losses: list[float] = []
    progress: list[tuple[int, float, float, float]] = []
    t0 = time.time()
    for epoch in range(epochs):                       # siren.py:98-105
        opt.zero_grad()
        pred = forward(params, omega_0, x)
        loss = F.mse_loss(pred, t_norm)
        loss.backward()
        opt.step()
        sched.step()
        losses.append(loss.item())
        if (epoch + 1) % log_every == 0:              # siren.py:107-115 (pre-step prediction)
            with torch.no_grad():
                real = pred * std + mean
                progress.append((epoch + 1, loss.item(), F.mse_loss(real, targets).item(),
                                 F.cosine_similarity(real, targets, dim=1).mean().item()))
    elapsed = time.time() - t0

    with torch.no_grad():
        m = final_metrics(forward(params, omega_0, x), targets, mean, std)
    state = {k: v.detach().cpu() for k, v in params.items()}
    n_params = count_parameters(state)
    raw = seq_len * d_head * 2                        # siren.py:127, fp16 KV baseline
    return OracleFit(
        state=state, target_mean=mean.cpu(), target_std=std.cpu(), losses=losses,
        final_mse=m['final_mse'], final_cosine_mean=m['final_cosine_mean'],
        final_cosine_min=m['final_cosine_min'], final_cosine_std=m['final_cosine_std'],
        per_pos_mse=m['per_pos_mse'].cpu().numpy(), cosine_sims=m['cosine_sims'].cpu().numpy(),
        compression_ratio=raw / (n_params * 4), raw_size_bytes=raw, siren_size_bytes=n_params * 4,
        num_parameters=n_params, train_time_seconds=elapsed, progress=progress,
    )


# --------------------------------------------------------------------------
# step-level pieces (known-answer tests for single kernels)
# --------------------------------------------------------------------------

def loss_and_grads(state: dict[str, torch.Tensor], omega_0: float, x: torch.Tensor,
                   t_norm: torch.Tensor) -> tuple[float, dict[str, torch.Tensor]]:
    """One forward + MSE + backward via autograd (siren.py:100-102)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in state.items()}
    loss = F.mse_loss(forward(params, omega_0, x), t_norm)
    loss.backward()
    return loss.item(), {k: v.grad.detach().clone() for k, v in params.items()}


def adam_reference_step(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor,
                        step: int, lr: float, beta1: float = 0.9, beta2: float = 0.999,
                        eps: float = 1e-8) -> None:
    """torch.optim.Adam single-tensor update, in place (torch/optim/adam.py
    `_single_tensor_adam`, non-capturable branch; defaults adam.py:38-42)."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


# --------------------------------------------------------------------------
# decode-side oracle (new functionality; the reference has no q.k kernel)
# --------------------------------------------------------------------------

def decode_scores(state: dict[str, torch.Tensor], omega_0: float, mean: torch.Tensor,
                  std: torch.Tensor, q: torch.Tensor, seq_len: int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """q . K_hat[n] for every cached position, with K_hat = SIREN(pos)*std+mean
    (reconstruction as in nerf_attention/evaluate.py:148-152).

    ``dtype=torch.float64`` evaluates the same fp32 weights and fp32 position grid in double precision: the
    exact function both fp32 evaluations (the reference's torch CPU arithmetic and the CUDA kernels)
    approximate.  With omega_0 = 30 per sine layer, fp32 rounding differences are amplified to a few 1e-5
    relative, so fp32 parity tests bound the kernel's error against this truth by the reference's own."""
    state = {k: v.to(dtype) for k, v in state.items()}
    k_hat = forward(state, omega_0, positions_for(seq_len).to(dtype)) * std.to(dtype) + mean.to(dtype)
    return k_hat @ q.to(dtype)


def kvread_scores(k_fp16: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """Plain KV-cache attention logits: fp16 keys from memory, fp32 accumulate."""
    return k_fp16.float() @ q.float()


def decode_attention(key_state: dict[str, torch.Tensor], value_state: dict[str, torch.Tensor], omega_k: float,
                     omega_v: float, mean_k: torch.Tensor, std_k: torch.Tensor, mean_v: torch.Tensor,
                     std_v: torch.Tensor, q: torch.Tensor, seq_len: int, scale: float,
                     dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """softmax(scale * q.K_hat) @ V_hat with both caches reconstructed from their SIRENs (the decode
    the reference describes in README.md:3-8; reconstruction as in evaluate.py:148-152).
    ``dtype=torch.float64``: the exact function of the same fp32 weights (see ``forward_exact``)."""
    pos = positions_for(seq_len).to(dtype)
    cast = lambda st: {k: v.to(dtype) for k, v in st.items()}
    k_hat = forward(cast(key_state), omega_k, pos) * std_k.to(dtype) + mean_k.to(dtype)
    v_hat = forward(cast(value_state), omega_v, pos) * std_v.to(dtype) + mean_v.to(dtype)
    p = torch.softmax(scale * (k_hat @ q.to(dtype)), dim=0)
    return p @ v_hat


def kvread_attention(k_fp16: torch.Tensor, v_fp16: torch.Tensor, q: torch.Tensor, scale: float) -> torch.Tensor:
    """Plain single-query attention over an fp16 KV cache, fp32 accumulate."""
    p = torch.softmax(scale * (k_fp16.float() @ q.float()), dim=0)
    return p @ v_fp16.float()
