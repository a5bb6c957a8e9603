"""Sweep driver: fit SIRENs to every (layer, head, key|value, architecture) job.

Same entry points, CLI flags, printed report and files as the reference
(nerf_attention/fit.py): ``fit_kv_cache``, ``fit_results.json`` records
(fit.py:95-118), ``{name}_model.pt`` checkpoints for the ``medium`` config
(fit.py:121-137).  What changed is underneath: the loop nest only *enumerates*
jobs; they are then trained together by nerf_attention.batched.fit_many, and
sharded over ranks when launched under torchrun (nerf_attention.sharding).
"""

from __future__ import annotations

import argparse
import json
from pathlib import Path

import numpy as np
import torch

from nerf_attention import sharding
from nerf_attention.batched import FitJob, fit_many
from nerf_attention.siren import SIREN
from nerf_attention.types import CONFIGS_FULL, CONFIGS_QUICK, FitResult, KVMetadata, SIRENConfig

RECORD_KEYS = ('name', 'layer', 'head', 'kv_type', 'config_name', 'hidden_features', 'hidden_layers',
               'omega_0', 'final_mse', 'final_cosine_mean', 'final_cosine_min', 'final_cosine_std',
               'compression_ratio', 'raw_size_bytes', 'siren_size_bytes', 'train_time_seconds',
               'num_parameters', 'seq_len', 'd_head')


def sweep_selection(metadata: KVMetadata, quick: bool) -> tuple[list[int], int, list[SIRENConfig]]:
    """Layers, heads per layer and architectures of a sweep (reference fit.py:39-48)."""
    nl = metadata.num_layers
    if quick:
        layers, heads, configs = [0, nl // 2, nl - 1], 1, CONFIGS_QUICK
    else:
        layers = [0, nl // 4, nl // 2, 3 * nl // 4, nl - 1]
        heads, configs = min(metadata.num_kv_heads, 4), CONFIGS_FULL
    return sorted({l for l in layers if l < nl}), heads, configs


def enumerate_job_specs(layers: list[int], heads: int, configs: list[SIRENConfig]) -> list[dict]:
    """Jobs in the reference's loop order: layer -> head -> key, value -> config (fit.py:54-65); no tensors yet."""
    return [{'name': f"L{layer_idx}_H{head_idx}_{kv_type}_{config.name}", 'layer': layer_idx, 'head': head_idx,
             'kv_type': kv_type, 'config': config, 'config_index': ci, 'tensor': None}
            for layer_idx in layers for head_idx in range(heads) for kv_type in ('key', 'value')
            for ci, config in enumerate(configs)]


def enumerate_jobs(layer_tensors: dict[int, dict[str, torch.Tensor]], layers: list[int], heads: int,
                   configs: list[SIRENConfig]) -> list[dict]:
    """The same with each job's [seq_len, d_head] tensor attached (layers without a file are skipped)."""
    jobs = enumerate_job_specs([l for l in layers if l in layer_tensors], heads, configs)
    for job in jobs:
        job['tensor'] = layer_tensors[job['layer']]['keys' if job['kv_type'] == 'key' else 'values'][job['head']]
    return jobs


def load_layers_async(kv_dir: Path, layers: list[int], pin: bool):
    """Start reading ``layer_XX.pt`` files (reference format, extract.py:159-162,241-244) on a thread pool and
    return {layer: Future[{'keys','values'}]}.  Files are memory-mapped and copied once into pinned host memory, so
    the upload in nerf_attention.batched is an asynchronous DMA; the caller builds its models meanwhile."""
    from concurrent.futures import ThreadPoolExecutor

    def load(path):
        try:
            blob = torch.load(path, map_location='cpu', weights_only=True, mmap=True)
        except (RuntimeError, ValueError):                # legacy (non-zip) serialisation cannot be mapped
            blob = torch.load(path, map_location='cpu', weights_only=True)
        out = {}
        for k in ('keys', 'values'):
            t = blob[k] if blob[k].dtype == torch.float32 else blob[k].float()
            out[k] = t.pin_memory() if pin else t
        return out

    if not layers:
        return {}
    pool = ThreadPoolExecutor(max_workers=min(8, len(layers)), thread_name_prefix='kv-load')
    futures = {l: pool.submit(load, Path(kv_dir) / f'layer_{l:02d}.pt') for l in layers}
    pool.shutdown(wait=False)
    return futures


def fit_kv_cache(
    kv_dir: Path,
    output_dir: Path,
    epochs: int = 5000,
    device: str = 'cuda',
    quick: bool = False,
    precision: str | None = None,
    seed_fn=None,
) -> list[dict]:
    """Fit SIRENs to an extracted KV cache and record metrics (reference fit.py:20-92).

    ``seed_fn(job) -> int | None`` (extension) seeds torch's CPU generator right before each
    model is constructed; the reference sets no seed, so the default is None.
    """
    kv_dir, output_dir = Path(kv_dir), Path(output_dir)
    output_dir.mkdir(parents=True, exist_ok=True)
    with open(kv_dir / 'metadata.json') as f:
        metadata = KVMetadata.from_dict(json.load(f))

    rank, _, world = sharding.env_world()
    distributed = sharding.ensure_process_group()
    chatty = rank == 0
    if chatty:
        print(f"KV Cache: {metadata.num_layers} layers x {metadata.num_kv_heads} heads")
        print(f"Per tensor: ({metadata.seq_len}, {metadata.head_dim}) = "
              f"{metadata.seq_len * metadata.head_dim * 2 / 1024:.1f} KB (float16 baseline)")
        print(f"Device: {device}, Epochs: {epochs}")

    layers, heads, configs = sweep_selection(metadata, quick)
    present = []
    for layer_idx in layers:
        if (kv_dir / f'layer_{layer_idx:02d}.pt').exists():
            present.append(layer_idx)
        elif chatty:
            print(f"  Skipping layer {layer_idx} (not found)")
    jobs = enumerate_job_specs(present, heads, configs)
    total = len(layers) * heads * 2 * len(configs)

    mine = list(range(len(jobs)))
    if distributed and world > 1:
        costs = [j['config'].flops_per_epoch(metadata.seq_len, metadata.head_dim) for j in jobs]
        keys = [(j['layer'], j['head'], j['kv_type']) for j in jobs]
        mine = sharding.shard_jobs(keys, costs, world)[rank]
    mine_set = set(mine)

    # each rank reads only the layer files its shard touches, in the background (pinned host memory)
    pending = load_layers_async(kv_dir, sorted({jobs[i]['layer'] for i in mine}),
                                pin=torch.device(device).type == 'cuda' and torch.cuda.is_available())

    # Model construction consumes torch's CPU generator in job order (reference fit.py:70 ->
    # siren.py:89).  Unseeded, every rank builds every model so that the stream -- and hence each
    # job's initial weights -- does not depend on the sharding; with per-job seeds only the
    # local jobs need building.  The files load meanwhile.
    models: dict[int, SIREN] = {}
    for i, job in enumerate(jobs):
        if seed_fn is not None:
            if i not in mine_set:
                continue
            seed = seed_fn(job)
            if seed is not None:
                torch.manual_seed(seed)
        models[i] = SIREN(job['config'], out_features=metadata.head_dim)

    layer_tensors = {l: f.result() for l, f in pending.items()}
    fit_jobs: dict[int, FitJob] = {}
    for i in mine:
        job = jobs[i]
        job['tensor'] = layer_tensors[job['layer']]['keys' if job['kv_type'] == 'key' else 'values'][job['head']]
        if job['tensor'].shape[1] != metadata.head_dim:   # metadata.json disagrees with the file: trust the file
            models[i] = SIREN(job['config'], out_features=job['tensor'].shape[1])
        fit_jobs[i] = FitJob(job['tensor'], job['config'], models[i], job['name'])

    # the fits train concurrently, so the reference's per-fit progress lines (siren.py:112-115) are
    # collected on the device and printed with each fit's record below
    results = fit_many([fit_jobs[i] for i in mine], epochs=epochs, device=device,
                       log_every=max(epochs // 5, 100), verbose=False, precision=precision,
                       progress=chatty and not (distributed and world > 1))
    progress_of = {jobs[i]['name']: getattr(r, 'progress', []) for i, r in zip(mine, results)}

    local_records = []
    for i, result in zip(mine, results):
        job = jobs[i]
        record = _result_to_record(job['name'], job['layer'], job['head'], job['kv_type'], result)
        local_records.append((i, record))
        if job['config'].name == 'medium':
            _save_model(output_dir, job['name'], result, record)

    if distributed and world > 1:
        all_records = _gather_records(jobs, local_records, device)
    else:
        all_records = [r for _, r in local_records]

    if chatty:
        for n, r in enumerate(all_records, 1):
            print(f"\n[{n}/{total}] {r['name']}")
            for e, norm_mse, real_mse, cos in progress_of.get(r['name'], []):
                print(f"  Epoch {e}/{epochs} | NormMSE: {norm_mse:.6f} | RealMSE: {real_mse:.6f} | CosSim: {cos:.4f}")
            print(f"  -> CosSim: {r['final_cosine_mean']:.4f} | "
                  f"Compress: {r['compression_ratio']:.1f}x | Time: {r['train_time_seconds']:.1f}s")
        with open(output_dir / 'fit_results.json', 'w') as f:
            json.dump(all_records, f, indent=2)
        _print_summary(all_records, layers)
    return all_records


def _gather_records(jobs: list[dict], local_records: list[tuple[int, dict]], device: str) -> list[dict]:
    """One all-gather of [job index, numeric fields...] rows; strings are rebuilt from the job list."""
    numeric = [k for k in RECORD_KEYS if k not in ('name', 'kv_type', 'config_name')]
    rows = np.array([[float(i)] + [float(r[k]) for k in numeric] for i, r in local_records],
                    dtype=np.float64).reshape(len(local_records), 1 + len(numeric))
    dev = device if (torch.distributed.get_backend() == 'nccl') else 'cpu'
    gathered = sharding.gather_rows(rows, dev)
    out = {}
    for row in gathered:
        i = int(row[0])
        rec = {'name': jobs[i]['name'], 'kv_type': jobs[i]['kv_type'], 'config_name': jobs[i]['config'].name}
        for k, v in zip(numeric, row[1:]):
            rec[k] = int(v) if k in ('layer', 'head', 'hidden_features', 'hidden_layers', 'raw_size_bytes',
                                     'siren_size_bytes', 'num_parameters', 'seq_len', 'd_head') else float(v)
        out[i] = {k: rec[k] for k in RECORD_KEYS}
    return [out[i] for i in sorted(out)]


def _result_to_record(name: str, layer: int, head: int, kv_type: str, result: FitResult) -> dict:
    """fit_results.json row; key set and order follow reference fit.py:98-118."""
    cfg = result.config
    values = (name, layer, head, kv_type, cfg.name, cfg.hidden_features, cfg.hidden_layers, cfg.omega_0,
              result.final_mse, result.final_cosine_mean, result.final_cosine_min, result.final_cosine_std,
              result.compression_ratio, result.raw_size_bytes, result.siren_size_bytes,
              result.train_time_seconds, result.num_parameters, result.seq_len, result.d_head)
    return dict(zip(RECORD_KEYS, values))


def detached_state(model: SIREN) -> dict[str, torch.Tensor]:
    """state_dict with every tensor copied out of its storage.  After a batched fit the parameters are views into
    one flat buffer holding ALL jobs' weights (batched.adopt_packed); torch.save serialises whole storages, so saving
    the views would write every fit's weights into every checkpoint."""
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def _save_model(output_dir: Path, name: str, result: FitResult, record: dict) -> None:
    """Checkpoint dict consumed by evaluate._load_model_from_checkpoint (reference fit.py:121-137)."""
    cfg = result.config
    torch.save({
        'model_state': detached_state(result.model),
        'config': {'hidden_features': cfg.hidden_features, 'hidden_layers': cfg.hidden_layers,
                   'omega_0': cfg.omega_0, 'name': cfg.name, 'out_features': result.d_head},
        'target_mean': result.target_mean,
        'target_std': result.target_std,
        'metrics': record,
    }, Path(output_dir) / f'{name}_model.pt')


def _print_summary(all_results: list[dict], layers_to_fit: list[int]) -> None:
    """Console report in the reference's layout (fit.py:140-180)."""
    bar = '=' * 80
    print(f"\n{bar}\nRESULTS SUMMARY\n{bar}")
    print(f"{'Name':<35} {'CosSim':>8} {'MSE':>10} {'Compress':>10} {'Time':>8}")
    print(f"{'-'*35} {'-'*8} {'-'*10} {'-'*10} {'-'*8}")
    for r in sorted(all_results, key=lambda x: x['final_cosine_mean'], reverse=True):
        print(f"{r['name']:<35} {r['final_cosine_mean']:>8.4f} {r['final_mse']:>10.6f} "
              f"{r['compression_ratio']:>9.1f}x {r['train_time_seconds']:>7.1f}s")
    print(f"\n{bar}\nKEY FINDINGS\n{bar}")

    def avg(rows, key):
        return float(np.mean([r[key] for r in rows]))

    for cn in sorted({r['config_name'] for r in all_results}):
        rows = [r for r in all_results if r['config_name'] == cn]
        print(f"  {cn:<10}: avg CosSim={avg(rows, 'final_cosine_mean'):.4f}, "
              f"avg Compression={avg(rows, 'compression_ratio'):.1f}x")
    key_rows = [r for r in all_results if r['kv_type'] == 'key']
    val_rows = [r for r in all_results if r['kv_type'] == 'value']
    if key_rows and val_rows:
        k_avg, v_avg = avg(key_rows, 'final_cosine_mean'), avg(val_rows, 'final_cosine_mean')
        print(f"\n  Keys avg CosSim:   {k_avg:.4f}")
        print(f"  Values avg CosSim: {v_avg:.4f}")
        gap = v_avg - k_avg
        print("  -> Values compress better (smoother signal)" if gap > 0.01 else
              "  -> Keys compress better (stronger positional structure)" if gap < -0.01 else
              "  -> Similar compressibility")
    for layer_idx in layers_to_fit:
        rows = [r for r in all_results if r['layer'] == layer_idx and r['config_name'] == 'medium']
        if rows:
            print(f"  Layer {layer_idx:2d} (medium): avg CosSim={avg(rows, 'final_cosine_mean'):.4f}")


def main() -> None:
    parser = argparse.ArgumentParser(description='Fit SIRENs to KV cache')
    parser.add_argument('--kv_dir', type=str, default='results/kv_cache')
    parser.add_argument('--output_dir', type=str, default='results/fits')
    parser.add_argument('--epochs', type=int, default=5000)
    parser.add_argument('--device', type=str, default='cuda')
    parser.add_argument('--quick', action='store_true')
    parser.add_argument('--precision', type=str, default=None, choices=['fp32', 'bf16'],
                        help="arithmetic of the H->H / H->D layers (default: $NERFATTN_PRECISION or fp32)")
    args = parser.parse_args()
    if args.device == 'cuda' and not torch.cuda.is_available():
        # the reference falls back to the CPU here (fit.py:192-194); this build has no CPU path
        raise SystemExit('CUDA not available: the B200 build of nerf_attention has no CPU fallback')
    fit_kv_cache(Path(args.kv_dir), Path(args.output_dir), args.epochs, args.device, args.quick,
                 precision=args.precision)


if __name__ == '__main__':
    main()
