"""Generate golden vectors by running the REAL reference (read-only, /root/reference).

Run in the build container only:  python tests/golden/make_golden.py
The reference cannot travel to the GPU box, so its outputs are committed here
(small .npz files) and replayed by tests/test_oracle_golden.py and the GPU
parity tests.  ``import nerf_attention`` of the reference pulls in matplotlib
(absent in this image) through analyze/evaluate, so the package __init__ is
bypassed with a stub module and only the torch/numpy submodules are imported
(SURVEY.md 8c).
"""

import hashlib
import json
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

REF = Path('/root/reference')
OUT = Path(__file__).resolve().parent


def import_reference():
    pkg = types.ModuleType('nerf_attention')
    pkg.__path__ = [str(REF / 'nerf_attention')]
    sys.modules['nerf_attention'] = pkg
    import nerf_attention.types as rtypes      # noqa: E402
    import nerf_attention.siren as rsiren      # noqa: E402
    import nerf_attention.extract as rextract  # noqa: E402
    return rtypes, rsiren, rextract


def state_digest(state) -> str:
    h = hashlib.sha256()
    for k in sorted(state):
        h.update(k.encode())
        h.update(state[k].detach().contiguous().numpy().tobytes())
    return h.hexdigest()


def synthetic_like(seed: int, n: int, d: int) -> torch.Tensor:
    """Small smooth-plus-noise tensor; saved verbatim in the fixture."""
    g = torch.Generator().manual_seed(seed)
    t = torch.linspace(0, 1, n).unsqueeze(1)
    f = torch.rand(1, d, generator=g) * 6 + 1
    ph = torch.rand(1, d, generator=g) * 6.28
    return (0.7 * torch.sin(6.2831853 * f * t + ph) + 0.2 * torch.randn(n, d, generator=g)
            + 0.3 * torch.rand(1, d, generator=g))


def main() -> None:
    torch.set_num_threads(1)   # deterministic reduction order for the fixtures
    rtypes, rsiren, rextract = import_reference()
    meta = {'torch': torch.__version__, 'cases': {}}

    # ---- 1. seeded init: digests + a slice of the numbers -------------------
    init_cases = {}
    for cfg in rtypes.CONFIGS_FULL:
        torch.manual_seed(1234)
        model = rsiren.SIREN(cfg, out_features=128)
        sd = model.state_dict()
        init_cases[cfg.name] = {
            'digest': state_digest(sd),
            'num_parameters': model.count_parameters(),
            'size_bytes': model.size_bytes(),
            'keys': list(sd.keys()),
            'w0_head': sd['network.0.linear.weight'].flatten()[:4].tolist(),
            'bf_tail': sd[f'network.{cfg.hidden_layers + 1}.bias'][-4:].tolist(),
        }
    meta['cases']['init_seed1234_d128'] = init_cases

    # ---- 2. forward known-answer -------------------------------------------
    arrays = {}
    for cfg in (rtypes.CONFIGS_FULL[0], rtypes.CONFIGS_FULL[2], rtypes.CONFIGS_FULL[5]):
        torch.manual_seed(77)
        model = rsiren.SIREN(cfg, out_features=128)
        pos = torch.linspace(0, 1, 96).unsqueeze(1)
        with torch.no_grad():
            arrays[f'fwd_{cfg.name}'] = model(pos).numpy()
    np.savez_compressed(OUT / 'forward.npz', **arrays)

    # ---- 3. full fit_siren trajectories ------------------------------------
    fit_specs = [
        # name, H, L, omega, N, D, epochs, data seed, model seed
        ('tiny', 64, 1, 30.0, 128, 16, 80, 3, 11),
        ('small', 128, 1, 30.0, 192, 32, 60, 4, 12),
        ('medium', 256, 2, 30.0, 128, 32, 40, 5, 13),
        ('hifreq', 256, 2, 60.0, 128, 16, 40, 6, 14),
        ('deep', 256, 3, 30.0, 64, 16, 30, 7, 15),
    ]
    fits = {}
    for name, h, l, w, n, d, epochs, dseed, mseed in fit_specs:
        cfg = rtypes.SIRENConfig(h, l, w, name)
        kv = synthetic_like(dseed, n, d)
        torch.manual_seed(mseed)
        res = rsiren.fit_siren(kv, cfg, epochs=epochs, lr=1e-4, device='cpu',
                               log_every=10 ** 9, verbose=False)
        fits[f'{name}_kv'] = kv.numpy()
        fits[f'{name}_losses'] = np.asarray(res.losses, dtype=np.float64)
        fits[f'{name}_cos'] = res.cosine_sims
        fits[f'{name}_ppmse'] = res.per_pos_mse
        fits[f'{name}_mean'] = res.target_mean.numpy()
        fits[f'{name}_std'] = res.target_std.numpy()
        fits[f'{name}_wf'] = res.model.state_dict()[f'network.{l + 1}.weight'].numpy()
        meta['cases'][f'fit_{name}'] = {
            'spec': [h, l, w, n, d, epochs, dseed, mseed],
            'final_mse': res.final_mse,
            'final_cosine_mean': res.final_cosine_mean,
            'final_cosine_min': res.final_cosine_min,
            'final_cosine_std': res.final_cosine_std,
            'compression_ratio': res.compression_ratio,
            'raw_size_bytes': res.raw_size_bytes,
            'siren_size_bytes': res.siren_size_bytes,
            'num_parameters': res.num_parameters,
            'state_digest': state_digest(res.model.state_dict()),
        }
    np.savez_compressed(OUT / 'fits.npz', **fits)

    # ---- 4. synthetic KV generator -----------------------------------------
    with tempfile.TemporaryDirectory() as tmp:
        rextract.extract_kv_cache_synthetic(seq_len=48, num_layers=3, num_kv_heads=2,
                                            head_dim=8, output_dir=Path(tmp))
        syn = {}
        for layer in range(3):
            blob = torch.load(Path(tmp) / f'layer_{layer:02d}.pt', weights_only=True)
            syn[f'keys_{layer}'] = blob['keys'].numpy()
            syn[f'values_{layer}'] = blob['values'].numpy()
        meta['cases']['synthetic_metadata'] = json.loads((Path(tmp) / 'metadata.json').read_text())
    np.savez_compressed(OUT / 'synthetic.npz', **syn)

    # ---- 5. config tables ---------------------------------------------------
    meta['cases']['configs_full'] = [[c.hidden_features, c.hidden_layers, c.omega_0, c.name]
                                     for c in rtypes.CONFIGS_FULL]
    meta['cases']['configs_quick'] = [[c.hidden_features, c.hidden_layers, c.omega_0, c.name]
                                      for c in rtypes.CONFIGS_QUICK]

    (OUT / 'golden.json').write_text(json.dumps(meta, indent=1))
    print('wrote', [p.name for p in OUT.glob('*.npz')], 'golden.json')


if __name__ == '__main__':
    main()
