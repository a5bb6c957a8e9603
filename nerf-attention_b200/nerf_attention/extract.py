"""Synthetic KV-cache generator (the input of every BASELINE config).

Bit-compatible with the reference generator (nerf_attention/extract.py:182-259):
one ``RandomState(layer * num_kv_heads + head)`` per (layer, head) consumed in the
same order -- per dimension: two base frequencies, the mid frequency, its phase,
then (position, width, amplitude) per spike, the key noise vector, the value
frequency and the value noise vector.  Real-model extraction (transformers +
bitsandbytes) is out of scope for this build.

Differences from the reference are organisational only: the per-(layer, head)
generator is exposed on its own so that a sweep can build just the layers it
fits, in memory, without a round trip through ``layer_XX.pt`` files.
"""

from __future__ import annotations

import argparse
import json
from pathlib import Path

import numpy as np
import torch

from nerf_attention.types import KVMetadata

_TWO_PI = 2 * np.pi


def synthetic_head(layer_idx: int, head_idx: int, seq_len: int, num_layers: int, num_kv_heads: int,
                   head_dim: int) -> tuple[torch.Tensor, torch.Tensor]:
    """Keys and values [seq_len, head_dim] fp32 of one (layer, head)."""
    rng = np.random.RandomState(layer_idx * num_kv_heads + head_idx)
    t = torch.linspace(0, 1, seq_len).numpy()            # fp32 grid: python_float * float32 array stays float32 (NumPy 2), as in the reference
    sharp = 1.0 + 2.0 * (layer_idx / max(num_layers - 1, 1))
    n_spikes = int(3 * sharp)
    max_width = max(2, int(5 / sharp))
    keys = np.empty((seq_len, head_dim), dtype=np.float32)
    values = np.empty((seq_len, head_dim), dtype=np.float32)
    for d in range(head_dim):
        f_lo = rng.uniform(1, 5)
        f_hi = rng.uniform(3, 10)
        smooth = 0.5 * np.sin(_TWO_PI * f_lo * t) + 0.3 * np.cos(_TWO_PI * f_hi * t)
        f_mid = rng.uniform(10, 30)
        ripple = 0.2 * np.sin(_TWO_PI * f_mid * t + rng.uniform(0, _TWO_PI))
        bumps = np.zeros(seq_len)
        for _ in range(n_spikes):
            centre = rng.randint(0, seq_len)
            width = rng.randint(1, max_width)
            amp = rng.uniform(0.5, 2.0)
            sigma = max(1, width / 2)
            lo, hi = max(0, centre - width), min(seq_len - 1, centre + width)
            for p in range(lo, hi + 1):
                bumps[p] += amp * np.exp(-0.5 * ((p - centre) / sigma) ** 2)
        key_noise = rng.randn(seq_len) * 0.1
        keys[:, d] = (smooth + ripple + bumps + key_noise).astype(np.float32)
        v_smooth = 0.6 * np.sin(_TWO_PI * rng.uniform(1, 8) * t)
        values[:, d] = (v_smooth + rng.randn(seq_len) * 0.15).astype(np.float32)
    return torch.from_numpy(keys), torch.from_numpy(values)


def stream_parameters(layer_idx: int, head_idx: int, num_layers: int, num_kv_heads: int) -> tuple[int, int, int]:
    """(seed, spikes per dimension, exclusive upper bound of the spike width) of one (layer, head) stream --
    reference extract.py:204,207,222-224."""
    sharp = 1.0 + 2.0 * (layer_idx / max(num_layers - 1, 1))
    return layer_idx * num_kv_heads + head_idx, int(3 * sharp), max(2, int(5 / sharp))


def synthetic_heads_cuda(pairs: list[tuple[int, int]], seq_len: int, num_layers: int, num_kv_heads: int,
                         head_dim: int, device='cuda') -> tuple[torch.Tensor, torch.Tensor]:
    """Keys and values [len(pairs), seq_len, head_dim] fp32 of the given (layer, head) pairs, generated on the
    GPU (csrc/synth.cuh through ``nerfattn_synth_kv``): the same numpy RandomState streams as ``synthetic_head``
    consumed in the same order, one CTA per stream.  The CPU generator costs ~0.1 s per head at N=2048 and
    ~5 min for the N=32768 data set; this is milliseconds per head.  Agreement with the CPU
    generator is to float32 rounding (see synth.cuh), not bit-for-bit."""
    import ctypes

    from nerf_attention import _native
    dev = _native.require_cuda(device)
    lib = _native.lib()
    n = len(pairs)
    with torch.cuda.device(dev):
        keys = torch.empty(n, seq_len, head_dim, device=dev)
        values = torch.empty(n, seq_len, head_dim, device=dev)
        positions = torch.linspace(0, 1, seq_len).to(dev)            # the CPU grid (extract.py:197), bit for bit
        streams = (_native.NaSynthStream * n)()
        for i, (layer, head) in enumerate(pairs):
            seed, n_spikes, max_width = stream_parameters(layer, head, num_layers, num_kv_heads)
            streams[i] = _native.NaSynthStream(seed, n_spikes, max_width, keys[i].data_ptr(), values[i].data_ptr())
        need = ctypes.c_size_t()
        _native.check(lib.nerfattn_synth_workspace_bytes(n, seq_len, head_dim, ctypes.byref(need)),
                      'nerfattn_synth_workspace_bytes')
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        _native.check(lib.nerfattn_synth_kv(streams, n, seq_len, head_dim, positions.data_ptr(), ws.data_ptr(),
                                            need.value, _native.stream_handle()), 'nerfattn_synth_kv')
        ws.record_stream(torch.cuda.current_stream())
        positions.record_stream(torch.cuda.current_stream())
    return keys, values


def synthetic_layer(layer_idx: int, seq_len: int, num_layers: int, num_kv_heads: int, head_dim: int,
                    heads: list[int] | None = None, device=None) -> dict[str, torch.Tensor]:
    """{'keys','values'}: [num_kv_heads, seq_len, head_dim]; heads not listed stay zero.
    ``device='cuda'`` generates on the GPU and returns device tensors."""
    wanted = list(range(num_kv_heads) if heads is None else heads)
    if device is not None and torch.device(device).type == 'cuda':
        keys = torch.zeros(num_kv_heads, seq_len, head_dim, device=device)
        values = torch.zeros(num_kv_heads, seq_len, head_dim, device=device)
        k, v = synthetic_heads_cuda([(layer_idx, h) for h in wanted], seq_len, num_layers, num_kv_heads,
                                    head_dim, device)
        keys[wanted], values[wanted] = k, v
        return {'keys': keys, 'values': values}
    keys = torch.zeros(num_kv_heads, seq_len, head_dim)
    values = torch.zeros(num_kv_heads, seq_len, head_dim)
    for h in wanted:
        keys[h], values[h] = synthetic_head(layer_idx, h, seq_len, num_layers, num_kv_heads, head_dim)
    return {'keys': keys, 'values': values}


def extract_kv_cache_synthetic(
    seq_len: int = 2048,
    num_layers: int = 32,
    num_kv_heads: int = 8,
    head_dim: int = 128,
    output_dir: Path = Path('results/kv_cache_synthetic'),
    layers: list[int] | None = None,
    device=None,
) -> KVMetadata:
    """Write layer_XX.pt + metadata.json like the reference.  ``layers`` (extension) limits
    the files written to the layers a sweep will actually read; ``device='cuda'`` (extension)
    generates on the GPU (``synthetic_heads_cuda``) and saves CPU tensors as the reference does."""
    print("Generating synthetic KV cache...")
    print(f"  {num_layers} layers, {num_kv_heads} heads, seq_len={seq_len}, head_dim={head_dim}")
    output_dir = Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    for layer_idx in (range(num_layers) if layers is None else layers):
        layer = synthetic_layer(layer_idx, seq_len, num_layers, num_kv_heads, head_dim, device=device)
        torch.save({k: v.cpu() for k, v in layer.items()}, output_dir / f'layer_{layer_idx:02d}.pt')
    metadata = KVMetadata(model_name='synthetic', num_layers=num_layers, num_kv_heads=num_kv_heads,
                          seq_len=seq_len, head_dim=head_dim, actual_tokens=seq_len)
    with open(output_dir / 'metadata.json', 'w') as f:
        json.dump(metadata.to_dict(), f, indent=2)
    total_mb = num_layers * num_kv_heads * seq_len * head_dim * 2 * 4 / 1024 / 1024
    print(f"Saved to {output_dir}/ ({total_mb:.1f} MB)")
    return metadata


def extract_kv_cache(*args, **kwargs):
    raise NotImplementedError(
        'real-model KV extraction (transformers + bitsandbytes 4-bit Llama) is outside the scope of '
        'the B200 hot-path build; use extract_kv_cache_synthetic or point --kv_dir at layer_XX.pt '
        'files produced by the reference extractor')


def main() -> None:
    parser = argparse.ArgumentParser(description='Extract KV cache')
    parser.add_argument('--model', type=str, default='meta-llama/Llama-3.1-8B')
    parser.add_argument('--seq_len', type=int, default=2048)
    parser.add_argument('--output_dir', type=str, default='results/kv_cache')
    parser.add_argument('--synthetic', action='store_true')
    parser.add_argument('--device', type=str, default='cuda')
    parser.add_argument('--synthetic_device', type=str, default='cpu',
                        help="'cuda': run the synthetic generator on the GPU (extension; default is the "
                             "reference's CPU generator)")
    args = parser.parse_args()
    if args.synthetic:
        extract_kv_cache_synthetic(seq_len=args.seq_len, output_dir=Path(args.output_dir + '_synthetic'),
                                   device=None if args.synthetic_device == 'cpu' else args.synthetic_device)
    else:
        extract_kv_cache(args.model, args.seq_len, Path(args.output_dir), args.device)


if __name__ == '__main__':
    main()
