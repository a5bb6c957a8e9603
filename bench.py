#!/usr/bin/env python
"""Headline benchmark: SIREN fit-epochs/s over the 280-fit sweep + decode us/token vs HBM KV read (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this framework, one JSON line
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

A *step* is one full pass of the hot path over the workload: the 280 fits of the sweep (5 layers x 4 KV heads x
key/value x 7 architectures, synthetic Llama-3.1-8B-shaped KV, N=2048, D=128) trained for `--epochs` (default
2000) full-batch Adam steps and evaluated.

  value      fits x epochs x steps / device time of the K timed steps, inputs resident in HBM (CUDA events, max over ranks)
  e2e        the same through the public API (nerf_attention.fit_many) from pinned host tensors, results read back
  scaling    "strong" (default): ONE 280-fit sweep sharded over the N ranks by (layer, head, key|value) unit with
             nerf_attention.sharding.shard_jobs -- no collective on the hot path, one all-gather of the final metrics;
             `weak_scaling` (N > 1) is the extra figure with a 280-fit sweep per rank
  roofline   the dominant kernel (chain::chain_kernel) timed alone, live, with CUDA events (profiling build of the library)
  decode     BASELINE's second metric: fused SIREN-eval + q.k vs a measured HBM KV-read + q.k, 512-32768 tokens
  quality    final CosSim of a 14-fit sample of the timed run against the oracle (bf16 gate 5e-3, fp32 gate 1e-3)
  fp32_mode  the parity mode's throughput (fewer epochs; stated) against the FP32 FMA pipe at the sampled clock
  cpu_baseline / torch_eager_b200   the reference's CPU path on the host cores, and its torch-eager path on this GPU
  config5    (8 GPUs, or --config5) 512 `medium` fits at 4096..32768 tokens sharded by (layer, head)
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT / 'nerf-attention_b200', ROOT):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import numpy as np   # noqa: E402
import torch         # noqa: E402

METRIC = 'siren_fit_epochs_per_sec_280fit_sweep'
UNIT = 'fit-epochs/s'
SWEEP_LAYERS = [0, 8, 16, 24, 31]       # reference fit.py:44-45 for 32 layers
SWEEP_HEADS = 4
NUM_LAYERS, NUM_KV_HEADS, HEAD_DIM = 32, 8, 128


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', choices=['native', 'reference'], default='native')
    ap.add_argument('--precision', choices=['bf16', 'fp32'], default='bf16')
    ap.add_argument('--scaling', choices=['strong', 'weak'], default='strong')
    ap.add_argument('--epochs', type=int, default=2000)
    ap.add_argument('--seq_len', type=int, default=2048)
    ap.add_argument('--no-e2e', action='store_true', help='device-resident timing and kernel phases only (profiling runs)')
    ap.add_argument('--no-extras', action='store_true', help='skip quality / fp32 / decode / CPU legs')
    ap.add_argument('--config5', action='store_true', help='also run BASELINE config 5 (default: only with 8 GPUs)')
    ap.add_argument('--cpu-epochs', type=int, default=300, help='epochs per architecture in the CPU sample')
    return ap.parse_args()


def world():
    return int(os.environ.get('RANK', 0)), int(os.environ.get('LOCAL_RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))


# ----------------------------------------------------------------------------- workload
def sweep_specs(rank: int, wsize: int, scaling: str, seq_len: int):
    """(layer, head, is_value, config index) of this rank's fits, in the reference's loop order (fit.py:54-65).
    strong: the one 280-fit sweep, sharded by (layer, head, key|value) unit, cost-balanced (sharding.shard_jobs --
    what `python -m nerf_attention.fit` does under torchrun); weak: a whole sweep per rank, layers shifted by rank."""
    import nerf_attention as na
    from nerf_attention.sharding import shard_jobs
    shift = rank if scaling == 'weak' else 0
    specs = [((base + shift) % NUM_LAYERS, head, is_value, ci)
             for base in SWEEP_LAYERS for head in range(SWEEP_HEADS) for is_value in (0, 1)
             for ci in range(len(na.CONFIGS_FULL))]
    if scaling == 'strong' and wsize > 1:
        costs = [na.CONFIGS_FULL[ci].flops_per_epoch(seq_len, HEAD_DIM) for _, _, _, ci in specs]
        keys = [(layer, head, is_value) for layer, head, is_value, _ in specs]
        specs = [specs[i] for i in shard_jobs(keys, costs, wsize)[rank]]
    return specs


def build_jobs(specs, seq_len: int, pin: bool):
    """FitJobs with the SURVEY 8d seed convention; each distinct tensor is generated (and pinned) once."""
    import nerf_attention as na
    from nerf_attention.extract import synthetic_head
    heads, tensors, jobs = {}, {}, []
    for layer, head, is_value, ci in specs:
        if (layer, head) not in heads:
            heads[(layer, head)] = synthetic_head(layer, head, seq_len, NUM_LAYERS, NUM_KV_HEADS, HEAD_DIM)
        key = (layer, head, is_value)
        if key not in tensors:
            t = heads[(layer, head)][is_value]
            tensors[key] = t.pin_memory() if pin else t
        cfg = na.CONFIGS_FULL[ci]
        torch.manual_seed(1000 * layer + 100 * head + 10 * is_value + ci)
        jobs.append(na.FitJob(tensors[key], cfg, na.SIREN(cfg, out_features=HEAD_DIM),
                              f"L{layer}_H{head}_{'value' if is_value else 'key'}_{cfg.name}"))
    return jobs


def workload_config(args, precision: str, wsize: int, fits_this_rank: int, reference_sample: str | None = None) -> dict:
    cfg = {'workload': f'sweep280 (BASELINE config 3): layers {SWEEP_LAYERS} x {SWEEP_HEADS} KV heads x key/value x 7 architectures '
                       f'(CONFIGS_FULL), synthetic Llama-3.1-8B-shaped KV [N={args.seq_len}, D={HEAD_DIM}], '
                       f'{args.epochs} epochs per fit; one step = the whole sweep',
           'fits': 280, 'fits_per_gpu': fits_this_rank, 'epochs': args.epochs, 'seq_len': args.seq_len, 'head_dim': HEAD_DIM,
           'precision': precision,
           'parallelism': (f'one 280-fit sweep sharded by (layer, head, key|value) unit over {wsize} GPU(s), no collective on the '
                           'hot path, one all-gather of the final metrics' if args.scaling == 'strong' else
                           f'{wsize} x 280 fits (a whole sweep per GPU)'),
           'l2': 'per-epoch working set (bf16 activations ~1.6 GB + Adam state 0.65 GB) >> 126 MB L2; no flush needed'}
    if reference_sample:
        cfg['workload'] = ('SAMPLE of sweep280 on the CPU -- ' + reference_sample + '; fit-epochs/s of the sample = the '
                           "sweep's rate on this CPU (equal fits per architecture); the full sweep would take hours")
        cfg['sampled'] = True
    return cfg


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
              'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc = gpu_index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.gpu_index}', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                 '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
        return self

    def stop(self) -> dict:
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, smax, power, reasons = [], [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, flag in zip(names, f[5:9]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        busy = [c for c, p in zip(sm, power) if p > 300] or sm
        return {'sm_mhz': statistics.median(busy) if busy else None, 'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None, 'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peaks() -> dict:
    path = ROOT / 'MEASURED_PEAKS.json'
    if path.exists():
        d = json.loads(path.read_text())
        return {'bf16_burst': d.get('bf16_tflops'), 'bf16_sustained': d.get('bf16_tflops_sustained'),
                'hbm_gbs': d.get('hbm_gbs'), 'source': 'MEASURED_PEAKS.json'}
    return {'bf16_burst': 1590.0, 'bf16_sustained': 1400.0, 'hbm_gbs': 6650.0, 'source': 'fallback (B200_PROFILING.md)'}


# ----------------------------------------------------------------------------- CPU / reference legs
def cpu_threads() -> int:
    # torchrun exports OMP_NUM_THREADS=1; the CPU leg is meant to use the box's host cores
    n = max(1, int(os.environ.get('NERFATTN_CPU_THREADS', os.cpu_count() or 1)))
    torch.set_num_threads(n)
    return n


def cpu_sample(seq_len: int, epochs: int, warm: int = 3) -> dict:
    """The reference's CPU path (oracle port of fit_siren) on this box's host cores: each of the 7 architectures on
    one synthetic key tensor for `epochs` epochs -- the sweep's own mix (equal fits per architecture), so
    7 * epochs / seconds is the sweep's fit-epochs/s on the CPU."""
    from nerf_attention.extract import synthetic_head
    from nerf_attention.types import CONFIGS_FULL
    from oracle import siren_oracle as orc
    threads = cpu_threads()
    kv, _ = synthetic_head(16, 0, seq_len, NUM_LAYERS, NUM_KV_HEADS, HEAD_DIM)
    per_cfg, total = {}, 0.0
    for ci, cfg in enumerate(CONFIGS_FULL):
        torch.manual_seed(16000 + ci)
        state = orc.init_state(cfg.hidden_features, cfg.hidden_layers, cfg.omega_0, HEAD_DIM)
        orc.fit(kv, cfg.hidden_features, cfg.hidden_layers, cfg.omega_0, epochs=warm, device='cpu',
                log_every=10 ** 9, init=state)
        t0 = time.perf_counter()
        orc.fit(kv, cfg.hidden_features, cfg.hidden_layers, cfg.omega_0, epochs=epochs, device='cpu',
                log_every=10 ** 9, init=state)
        dt = time.perf_counter() - t0
        per_cfg[cfg.name] = epochs / dt
        total += dt
    return {'value': 7 * epochs / total, 'unit': UNIT, 'cores': threads, 'kind': 'port',
            'host_cpus': os.cpu_count(), 'seconds': total, 'epochs_per_sec_by_arch': per_cfg,
            'sample': f'oracle port of reference fit_siren (torch CPU, {threads} threads): 7 architectures x '
                      f'1 synthetic key tensor [{seq_len}x{HEAD_DIM}] x {epochs} epochs after {warm} warm-up epochs'}


def cpu_config1(epochs: int = 2000) -> dict:
    """BASELINE config 1 in full (reference quickstart.py:34-58, `uv run quickstart --cpu`): synthetic KV with 512
    tokens, 4 layers x 4 heads; fit_kv_cache(quick=True) = layers {0, 2, 3} x head 0 x key/value x {small, medium}
    = 12 fits x 2000 epochs on the CPU, one after the other as the reference runs them."""
    from nerf_attention.extract import synthetic_head
    from nerf_attention.types import CONFIGS_QUICK
    from oracle import siren_oracle as orc
    threads = cpu_threads()
    t0, fits = time.perf_counter(), 0
    cos = []
    for layer in (0, 2, 3):
        keys, values = synthetic_head(layer, 0, 512, 4, 4, HEAD_DIM)
        for kv in (keys, values):
            for cfg in CONFIGS_QUICK:
                r = orc.fit(kv, cfg.hidden_features, cfg.hidden_layers, cfg.omega_0, epochs=epochs, device='cpu',
                            log_every=10 ** 9)
                cos.append(r.final_cosine_mean)
                fits += 1
    dt = time.perf_counter() - t0
    return {'fits': fits, 'epochs': epochs, 'seq_len': 512, 'seconds': dt, 'fit_epochs_per_sec': fits * epochs / dt,
            'cores': threads, 'mean_final_cossim': float(np.mean(cos)),
            'what': 'BASELINE config 1 (reference quickstart --cpu, fit step): 12 fits x 2000 epochs, oracle port on the host cores'}


def run_reference(args) -> None:
    rank, _, wsize = world()
    if rank != 0:
        return
    for _ in range(max(args.warmup, 0) and 1):
        cpu_sample(args.seq_len, max(2, args.cpu_epochs // 8), warm=1)
    vals, secs, last = [], [], None
    for _ in range(max(args.steps, 1)):
        last = cpu_sample(args.seq_len, args.cpu_epochs)
        vals.append(last['value']); secs.append(last['seconds'])
    value = 7 * args.cpu_epochs * len(vals) / sum(secs)
    last['value'] = value
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * sum(secs) / len(secs),
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(args, 'fp32', wsize, 280, reference_sample=last['sample']),
        'cpu_baseline': last,
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    if not args.no_extras:
        line['config1_cpu'] = cpu_config1()
    print(json.dumps(line))


def torch_eager_on_gpu(seq_len: int, epochs: int = 60) -> dict:
    """The reference's own torch-eager path on this B200 (oracle port, device='cuda'), medium config."""
    from nerf_attention.extract import synthetic_head
    from oracle import siren_oracle as orc
    kv, _ = synthetic_head(16, 0, seq_len, NUM_LAYERS, NUM_KV_HEADS, HEAD_DIM)
    torch.manual_seed(1)
    state = orc.init_state(256, 2, 30.0, HEAD_DIM)
    orc.fit(kv, 256, 2, 30.0, epochs=10, device='cuda', log_every=10 ** 9, init=state)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    orc.fit(kv, 256, 2, 30.0, epochs=epochs, device='cuda', log_every=10 ** 9, init=state)
    torch.cuda.synchronize()
    return {'medium_epochs_per_sec': epochs / (time.perf_counter() - t0),
            'sample': f'oracle port on cuda, medium, [{seq_len}x{HEAD_DIM}], {epochs} epochs, one fit at a time'}


# ----------------------------------------------------------------------------- extra records of the native arm
def quality_record(jobs, states, results, args, dev) -> dict:
    """Final CosSim of a 14-fit sample of the timed run (2 units x 7 architectures, the run's own initial weights)
    against the oracle: torch eager fp32 on this GPU (the reference's default device path, TF32 off), same epochs."""
    import nerf_attention as na
    from oracle import siren_oracle as orc
    idx = list(range(min(14, len(jobs))))
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    t0 = time.perf_counter()
    try:
        ref = [orc.fit(jobs[i].kv_tensor, jobs[i].config.hidden_features, jobs[i].config.hidden_layers,
                       jobs[i].config.omega_0, epochs=args.epochs, device=str(dev), log_every=10 ** 9,
                       init={k: v.clone() for k, v in states[i].items()}) for i in idx]
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    rec = {'sample': [jobs[i].name for i in idx], 'epochs': args.epochs, 'oracle': 'oracle port, torch eager fp32 on cuda, TF32 off',
           'oracle_seconds': time.perf_counter() - t0, 'gate': {'bf16': 5e-3, 'fp32': 1e-3}}
    diffs = [abs(results[i].final_cosine_mean - ref[k].final_cosine_mean) for k, i in enumerate(idx)]
    rec[f'max_abs_cos_diff_{args.precision}'] = max(diffs)
    other = 'fp32' if args.precision == 'bf16' else 'bf16'
    ojobs = []
    for i in idx:
        m = na.SIREN(jobs[i].config, out_features=HEAD_DIM)
        m.load_state_dict(states[i])
        ojobs.append(na.FitJob(jobs[i].kv_tensor, jobs[i].config, m, jobs[i].name))
    ores = na.fit_many(ojobs, epochs=args.epochs, device=str(dev), verbose=False, precision=other)
    rec[f'max_abs_cos_diff_{other}'] = max(abs(r.final_cosine_mean - ref[k].final_cosine_mean) for k, r in enumerate(ores))
    rec['max_abs_cos_diff'] = {'bf16': rec['max_abs_cos_diff_bf16'], 'fp32': rec['max_abs_cos_diff_fp32']}
    rec['within_gate'] = bool(rec['max_abs_cos_diff_bf16'] <= 5e-3 and rec['max_abs_cos_diff_fp32'] <= 1e-3)
    return rec


def fp32_record(jobs, initial, args, dev, clocks_mhz) -> dict:
    """The parity mode (SIMT fp32 FMA everywhere) on the same 280 fits: throughput is independent of the epoch count,
    so it is timed over fewer epochs (stated); roofline = the FP32 FMA pipe at the clock sampled during the run."""
    from nerf_attention import batched
    epochs = max(10, min(200, args.epochs))
    for j, flat in zip(jobs, initial):
        batched.adopt_packed(j.model, flat)
    batch = batched.FitBatch(jobs, epochs=epochs, device=str(dev), precision='fp32', keep_initial=True)
    batch.launch()                                            # warm-up
    torch.cuda.synchronize()
    sampler = ClockSampler(dev.index or 0).start()
    batch.reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); batch.launch(); e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    sec = e0.elapsed_time(e1) / 1e3
    batch.collect()
    flops = sum(j.config.flops_per_epoch(args.seq_len, HEAD_DIM) for j in jobs) * epochs
    mhz = clocks.get('sm_mhz') or clocks_mhz or 1965.0
    peak = 148 * 128 * 2 * mhz * 1e6 / 1e12
    return {'value': len(jobs) * epochs / sec, 'unit': UNIT, 'epochs': epochs, 'fits': len(jobs), 'seconds': sec,
            'tflops': flops / sec / 1e12, 'roofline': {'bound': 'fp32-fma', 'achieved': flops / sec / 1e12, 'peak': peak,
                                                       'unit': 'TFLOP/s', 'frac': flops / sec / 1e12 / peak,
                                                       'peak_source': f'148 SM x 128 lanes x 2 x {mhz:.0f} MHz (median SM clock during this run)'},
            'clocks': clocks, 'kernel': 'f32::sgemm_kernel family (fwd-sine / fwd-out / dX / dW split-K) + adam_kernel',
            'note': f'NA_PREC_FP32, {epochs} epochs per fit (throughput does not depend on the epoch count)'}


def decode_record(dev, peaks) -> dict:
    """BASELINE's second metric (config 4 + the 32768-token end of the crossover table): fused SIREN-eval + q.k (bf16
    tensor path, K never materialised) vs a measured fp16 KV-read + q.k from HBM, `medium` architecture.  Heads per
    launch are chosen so that the KV side streams >= 256 MB per launch (a 256-head launch at 512 tokens is 33 MB = 5 us
    at HBM speed, which measures launch ramp, not bandwidth); both sides process the same heads per launch."""
    import nerf_attention as na
    from nerf_attention.evaluate import profile_decode
    cfg = next(c for c in na.CONFIGS_FULL if c.name == 'medium')
    torch.manual_seed(0)
    models = [na.SIREN(cfg, out_features=HEAD_DIM) for _ in range(8)]
    sampler = ClockSampler(dev.index or 0).start()
    rows = []
    for n in (512, 1024, 2048, 4096, 32768):
        heads = int(min(2048, max(64, (256 << 20) // (n * HEAD_DIM * 2))))
        r = profile_decode(models, [n], heads_per_launch=heads, precisions=('bf16',), device=str(dev), warmup=5, runs=30)[0]
        rows.append({'seq_len': n, 'heads_per_launch': heads, 'kv_bytes_per_launch': r['kv_bytes_per_launch'],
                     'kvread_us': r['kvread_us'], 'kvread_gbs': r['kvread_gbs'], 'kvread_frac_of_hbm_peak': r['kvread_gbs'] / peaks['hbm_gbs'],
                     'kvread_us_per_token_head': r['kvread_us_per_token_head'],
                     'siren_bf16_us': r['siren_bf16_us'], 'siren_bf16_tflops': r['siren_bf16_tflops'],
                     'siren_bf16_frac_of_bf16_peak': r['siren_bf16_tflops'] / peaks['bf16_sustained'],
                     'siren_bf16_us_per_token_head': r['siren_bf16_us_per_token_head'],
                     'siren_over_kvread': r['siren_bf16_over_kvread']})
    return {'metric': 'decode us per token-head: fused SIREN-eval + q.k vs HBM KV-read + q.k', 'architecture': 'medium (256 x 2, omega_0 30)',
            'rows': rows, 'clocks': sampler.stop(), 'hbm_peak_gbs': peaks['hbm_gbs'], 'bf16_peak_tflops': peaks['bf16_sustained'],
            'crossover': 'none in 512-32768 tokens: the HBM read is faster at every length (the reference reaches the same conclusion, FINDINGS.md)',
            'kernels': 'chain::chain_kernel<256,2,1,1> + dec::decode_prep/finish (SIREN side), dec::kvread_qk_kernel (KV side)'}


def config5_record(rank, wsize, dev, dist, seq_lens=(4096, 8192, 16384, 32768), epochs=40) -> dict | None:
    """BASELINE config 5: `medium` fits for all 32 layers x 8 KV heads x key/value = 512 fits, sharded by (layer, head)
    over the ranks, at 4096..32768 tokens; inputs come from the GPU synthetic generator (nerfattn_synth_kv)."""
    import nerf_attention as na
    from nerf_attention import batched
    from nerf_attention.sharding import shard_jobs
    cfg = next(c for c in na.CONFIGS_FULL if c.name == 'medium')
    units = [(layer, head) for layer in range(NUM_LAYERS) for head in range(NUM_KV_HEADS)]
    mine = [units[i] for i in shard_jobs(units, [1.0] * len(units), wsize)[rank]]
    rows = []
    for n in seq_lens:
        kv = config5_inputs(mine, n, dev)                     # {(layer, head): (keys, values)} on the device
        jobs = []
        for (layer, head) in mine:
            for is_value in (0, 1):
                torch.manual_seed(1000 * layer + 100 * head + 10 * is_value + 2)
                jobs.append(na.FitJob(kv[(layer, head)][is_value], cfg, na.SIREN(cfg, out_features=HEAD_DIM)))
        batch = batched.FitBatch(jobs, epochs=epochs, device=str(dev), precision='bf16', keep_initial=True)
        batch.launch()
        torch.cuda.synchronize()
        if wsize > 1:
            dist.barrier()
        batch.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); batch.launch(); e1.record()
        torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3
        if wsize > 1:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        res = batch.collect()
        cosm = float(np.mean([r.final_cosine_mean for r in res]))
        flops = 512 * cfg.flops_per_epoch(n, HEAD_DIM) * epochs
        rows.append({'seq_len': n, 'fits': 512, 'fits_per_gpu': len(jobs), 'epochs': epochs, 'seconds': sec,
                     'fit_epochs_per_sec': 512 * epochs / sec, 'tflops_total': flops / sec / 1e12,
                     'tflops_per_gpu': flops / sec / 1e12 / wsize, 'mean_cossim_after_epochs': cosm})
        del batch, jobs, kv
        torch.cuda.empty_cache()
    return {'workload': 'BASELINE config 5: 512 medium fits (32 layers x 8 KV heads x key/value), sharded by (layer, head)',
            'n_gpus': wsize, 'rows': rows,
            'note': f'{epochs} epochs per fit (throughput measurement; the reference trains 2000), inputs generated on the GPU'}


def config5_inputs(units, seq_len: int, dev) -> dict:
    """Synthetic keys / values of the given (layer, head) units on the device (csrc/synth.cuh, the GPU generator)."""
    from nerf_attention.extract import synthetic_heads_cuda
    keys, values = synthetic_heads_cuda(list(units), seq_len, NUM_LAYERS, NUM_KV_HEADS, HEAD_DIM, device=str(dev))
    return {u: (keys[k], values[k]) for k, u in enumerate(units)}


# ----------------------------------------------------------------------------- native arm
def run_native(args) -> None:
    import torch.distributed as dist
    import nerf_attention as na
    from nerf_attention import _native, batched
    from nerf_attention.sharding import gather_rows

    rank, local_rank, wsize = world()
    assert torch.cuda.is_available(), 'bench.py needs a B200; there is no CPU fallback'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if wsize > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL announces its version on stdout when the first communicator is built; stdout carries exactly one JSON
        # line, so the set-up and the first collective run with fd 1 pointed at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', rank=rank, world_size=wsize, device_id=dev)
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    _native.lib()                                            # fail loudly before any timing

    def barrier():
        if wsize > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(seconds: float) -> float:
        if wsize == 1:
            return seconds
        t = torch.tensor([seconds], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_sweep(jobs, initial, steps, warmup, sample_clocks):
        """Device-resident arm: inputs and initial weights already in HBM; every step restores them, trains, and (N > 1)
        all-gathers the final scalars -- the sweep's one collective."""
        for j, flat in zip(jobs, initial):
            batched.adopt_packed(j.model, flat)
        batch = batched.FitBatch(jobs, epochs=args.epochs, device=str(dev), precision=args.precision, keep_initial=True)
        scal = batch.scal.buf
        pad = None
        if wsize > 1:                                         # ranks own different numbers of fits: pad to the largest
            n_max = torch.tensor([scal.numel()], device=dev)
            dist.all_reduce(n_max, op=dist.ReduceOp.MAX)
            pad = torch.zeros(int(n_max.item()), device=dev)

        def step():
            batch.reset()
            batch.launch()
            if wsize > 1:
                pad[:scal.numel()].copy_(scal)
                dist.all_gather([torch.empty_like(pad) for _ in range(wsize)], pad)
        for _ in range(warmup):
            step()
        barrier()
        sampler = ClockSampler(local_rank).start() if (sample_clocks and rank == 0) else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        clocks = sampler.stop() if sampler else None
        seconds = max_over_ranks(e0.elapsed_time(e1) / 1e3)
        return batch, seconds, clocks

    specs = sweep_specs(rank, wsize, args.scaling, args.seq_len)
    jobs = build_jobs(specs, args.seq_len, pin=True)
    states = [{k: v.clone() for k, v in j.model.state_dict().items()} for j in jobs[:14]]     # for the quality sample
    initial = []                                             # seeded initial weights of every job, on the host
    for j in jobs:
        flat = torch.empty(j.model.count_parameters(), dtype=torch.float32).pin_memory()
        batched.pack_model(j.model, flat)
        initial.append(flat)
    n_fits_total = 280 if args.scaling == 'strong' else 280 * wsize
    fit_epochs_per_step = n_fits_total * args.epochs
    my_flops = sum(j.config.flops_per_epoch(args.seq_len, HEAD_DIM) for j in jobs) * args.epochs
    flops_all = torch.tensor([float(my_flops)], dtype=torch.float64, device=dev)
    if wsize > 1:
        dist.all_reduce(flops_all)
    total_flops = float(flops_all.item())                   # all ranks, one step

    batch, dev_seconds, clocks = timed_sweep(jobs, initial, args.steps, args.warmup, True)
    results = batch.collect()
    launches = batch.launches_per_call() * args.steps
    del batch
    torch.cuda.empty_cache()

    # ---- the kernels of a step, one class at a time, live: the profiling build of the library (libnerfattn_prof.so)
    # honours NERFATTN_PHASE (1 = chain kernels only, 2 = dW + Adam kernels only, 16 = the fit-resident kernels of the
    # tiny / small fits only -- they run beside the others in a real step --, 8 = none; the 2H-parameter
    # layer-0 Adam launch that ends an epoch always runs), so the difference to the "neither" run is that class's
    # device time per epoch (CUDA events on this stream; results of such runs are meaningless and discarded).
    phases = None
    if rank == 0 and args.precision == 'bf16' and os.environ.get('NERFATTN_NO_CHAIN', '0') in ('', '0'):
        pe = max(20, min(100, args.epochs))
        pbatch = batched.FitBatch(jobs, epochs=pe, device=str(dev), precision=args.precision, keep_initial=True,
                                  lib=_native.prof_lib())

        def phase_ms(mask: int) -> float:
            os.environ['NERFATTN_PHASE'] = str(mask)
            try:
                best = None
                for _ in range(3):
                    pbatch.reset()
                    torch.cuda.synchronize()
                    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    p0.record(); pbatch.launch(); p1.record()
                    torch.cuda.synchronize()
                    ms = p0.elapsed_time(p1)
                    best = ms if best is None else min(best, ms)
                return best
            finally:
                os.environ.pop('NERFATTN_PHASE', None)
        base_ms = phase_ms(8)
        phases = {'epochs': pe, 'fixed_ms': base_ms,
                  'chain_ms_per_epoch': (phase_ms(1) - base_ms) / pe,
                  'dw_adam_ms_per_epoch': (phase_ms(2) - base_ms) / pe,
                  'resident_ms_per_epoch': (phase_ms(16) - base_ms) / pe,
                  'all_ms_per_epoch': (phase_ms(7) - base_ms) / pe,
                  'note': 'fixed_ms = set-up + final evaluation + the layer-0 Adam launches of all epochs'}
        pbatch.collect()
        del pbatch
        torch.cuda.empty_cache()
    # mean final CosSim of the whole sweep (all ranks): [is_value, cos] rows through the metrics all-gather
    cos_rows = gather_rows(np.array([[s[2], r.final_cosine_mean] for s, r in zip(specs, results)], dtype=np.float64), dev)
    cos_keys = float(cos_rows[cos_rows[:, 0] == 0, 1].mean())
    cos_vals = float(cos_rows[cos_rows[:, 0] == 1, 1].mean())

    # ---- end-to-end arm: public API, host tensors in, results out, every step
    e2e = None
    if not args.no_e2e:
        def gather_metrics(res) -> np.ndarray:
            rows = np.array([[*s, r.final_cosine_mean, r.final_mse] for s, r in zip(specs, res)], dtype=np.float64)
            return gather_rows(rows, dev)

        def e2e_step(rebuild: bool = False):
            if rebuild:                                       # variant: seeded CPU model construction counted as well
                for j, (layer, head, is_value, ci) in zip(jobs, specs):
                    torch.manual_seed(1000 * layer + 100 * head + 10 * is_value + ci)
                    j.model = na.SIREN(na.CONFIGS_FULL[ci], out_features=HEAD_DIM)
            else:
                for j, flat in zip(jobs, initial):
                    batched.adopt_packed(j.model, flat)       # host views of the initial weights
            res = na.fit_many(jobs, epochs=args.epochs, device=str(dev), verbose=False, precision=args.precision)
            gather_metrics(res)
            return batched.last_stats
        e2e_step()                                            # warm-up (allocator, pinned pools)
        barrier()
        t0 = time.perf_counter()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        stats = None
        for _ in range(args.steps):
            stats = e2e_step()
        ev1.record()
        barrier()
        e2e_seconds = max_over_ranks(max(time.perf_counter() - t0, ev0.elapsed_time(ev1) / 1e3))
        barrier()
        t1 = time.perf_counter()
        e2e_step(rebuild=True)
        barrier()
        rebuild_seconds = max_over_ranks(time.perf_counter() - t1)
        e2e = {'value': fit_epochs_per_step * args.steps / e2e_seconds, 'unit': UNIT,
               'h2d_bytes_per_step': stats.h2d_bytes, 'd2h_bytes_per_step': stats.d2h_bytes,
               'ms_per_step': 1e3 * e2e_seconds / args.steps,
               'host_setup_ms_per_step': 1e3 * stats.setup_seconds,
               'api': 'nerf_attention.fit_many(this rank\'s FitJobs: pinned host KV tensors + pre-built seeded models whose weights '
                      'are host tensors) + metrics all-gather; every step packs and uploads tensors and weights (H2D) and '
                      'reads losses/metrics back (D2H); byte counts are this rank\'s',
               'value_including_cpu_model_construction': fit_epochs_per_step / rebuild_seconds,
               'note': 'the second value also counts building the seeded nn.Module SIRENs on the CPU (torch CPU RNG), '
                       'as the reference does inside fit_siren (siren.py:89); one step'}

    # ---- weak-scaling figure (N > 1): a whole 280-fit sweep per rank, one timed step
    weak = None
    if wsize > 1 and args.scaling == 'strong' and not args.no_extras:
        wjobs = build_jobs(sweep_specs(rank, wsize, 'weak', args.seq_len), args.seq_len, pin=False)
        winit = []
        for j in wjobs:
            flat = torch.empty(j.model.count_parameters(), dtype=torch.float32)
            batched.pack_model(j.model, flat)
            winit.append(flat)
        wb, wsec, _ = timed_sweep(wjobs, winit, 1, 1, False)
        wb.collect()
        del wb, wjobs
        torch.cuda.empty_cache()
        weak = {'value': 280 * wsize * args.epochs / wsec, 'unit': UNIT, 'fits_per_gpu': 280, 'steps': 1, 'warmup': 1,
                'ms_per_step': 1e3 * wsec, 'what': f'{wsize} x 280 fits: every rank trains a whole sweep on its own layers'}

    config5 = None
    if (args.config5 or wsize == 8) and not args.no_extras:
        config5 = config5_record(rank, wsize, dev, dist)

    if rank != 0:
        if wsize > 1:
            dist.destroy_process_group()
        return

    peaks = measured_peaks()
    achieved = total_flops * args.steps / dev_seconds / 1e12 / wsize       # per GPU: max-over-ranks time
    traffic = None
    tpath = ROOT / 'profiles' / 'ncu_traffic.json'
    if tpath.exists():
        traffic = json.loads(tpath.read_text())
    if args.precision == 'bf16' and phases:
        # chain kernel = forward + dX of every fit: 4N(LH^2+HD) + 2NH FLOPs per fit-epoch (dW is the rest of F)
        chain_flops = sum(4 * args.seq_len * (j.config.hidden_layers * j.config.hidden_features ** 2 +
                                              j.config.hidden_features * HEAD_DIM) +
                          2 * args.seq_len * j.config.hidden_features for j in jobs)
        k_ach = chain_flops / (phases['chain_ms_per_epoch'] * 1e-3) / 1e12
        n_params = sum(j.model.count_parameters() for j in jobs)
        act_bytes = sum(((2 * (j.config.hidden_layers + 1) - 1) * args.seq_len * j.config.hidden_features +
                         args.seq_len * HEAD_DIM) * 2 for j in jobs)     # bf16 h_l, dz_l (l >= 1), dY: read once by dW
        upd_bytes = act_bytes + 26 * n_params                           # + p, m, v read and written (24 B) + bf16 mirror (2 B)
        upd_gbs = upd_bytes / (phases['dw_adam_ms_per_epoch'] * 1e-3) / 1e9
        roof = {'bound': 'tensor', 'achieved': k_ach, 'peak': peaks['bf16_sustained'], 'unit': 'TFLOP/s',
                'frac': k_ach / peaks['bf16_sustained'],
                'traffic': (traffic or {}).get('chain_dram_bytes_per_epoch'),
                'peak_source': f"{peaks['source']} bf16_tflops_sustained (kernels timed inside a seconds-long step)",
                'kernel': 'chain::chain_kernel<H,NS,0,CL> (fused forward + loss + dX + layer-0 gradient per 128-row tile; one launch '
                          'per shape group and epoch): algorithmic FLOPs 4N(LH^2+HD)+2NH per fit-epoch, all fits of this GPU / '
                          'summed device time of the launches of one epoch (NERFATTN_PHASE=1 minus =8 in libnerfattn_prof.so, CUDA events)',
                'per_launch': 'one epoch = one chain launch per shape group (5 for the sweep); achieved/traffic are per epoch (sum over them)',
                'share_of_step': phases['chain_ms_per_epoch'] / phases['all_ms_per_epoch'],
                'phases_ms_per_epoch': phases,
                'dw_adam_kernel': {'bound': 'hbm', 'achieved': upd_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                   'frac': upd_gbs / peaks['hbm_gbs'],
                                   'traffic': (traffic or {}).get('dw_adam_dram_bytes_per_epoch'),
                                   'algorithmic_bytes_per_epoch': upd_bytes,
                                   'kernel': 'dw::dw_adam_kernel<BN> (grouped dW GEMMs of all layers + Adam, one launch per shape group '
                                             'and epoch): bf16 operands read once + 26 B per parameter'},
                'whole_step': {'achieved': achieved, 'frac': achieved / peaks['bf16_sustained'],
                               'note': 'all kernels: F = 6N(LH^2+HD)+4NH per fit-epoch / step time',
                               'traffic': (traffic or {}).get('step_dram_bytes_per_epoch')}}
    elif args.precision == 'bf16':
        roof = {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['bf16_sustained'], 'unit': 'TFLOP/s',
                'frac': achieved / peaks['bf16_sustained'], 'traffic': None,
                'peak_source': f"{peaks['source']} bf16_tflops_sustained (kernels timed inside a seconds-long step)",
                'kernel': 'whole step (unfused tc_gemm_kernel path, NERFATTN_NO_CHAIN=1)'}
    else:
        fp32_peak = 148 * 128 * 2 * (clocks['sm_mhz'] or 1965.0) * 1e6 / 1e12
        roof = {'bound': 'fp32-fma', 'achieved': achieved, 'peak': fp32_peak, 'unit': 'TFLOP/s',
                'frac': achieved / fp32_peak, 'traffic': None,
                'peak_source': '148 SM x 128 lanes x 2 x median SM clock under load',
                'kernel': 'whole step (sgemm_kernel family)'}

    line = {
        'metric': METRIC, 'value': fit_epochs_per_step * args.steps / dev_seconds, 'unit': UNIT,
        'n_gpus': wsize, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dev_seconds / args.steps,
        'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'bf16' if args.precision == 'bf16' else 'f32', 'data': 'synthetic',
        'config': workload_config(args, args.precision, wsize, len(jobs)),
        'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches, 'roofline': roof,
        'quality': {'cos_keys_mean': cos_keys, 'cos_values_mean': cos_vals},
        'tflops_per_gpu': achieved,
    }
    if weak:
        line['weak_scaling'] = weak
    if config5:
        line['config5'] = config5
    if not args.no_e2e and not args.no_extras:
        try:
            line['quality'].update(quality_record(jobs, states, results, args, dev))
        except Exception as exc:                              # keep the line; the tests assert the same gates
            line['quality']['error'] = repr(exc)
        if wsize == 1:
            try:
                if args.precision == 'bf16':
                    line['fp32_mode'] = fp32_record(jobs, initial, args, dev, (clocks or {}).get('sm_mhz'))
                line['decode'] = decode_record(dev, peaks)
            except Exception as exc:
                line['extras_error'] = repr(exc)
            line['cpu_baseline'] = cpu_sample(args.seq_len, args.cpu_epochs)
            try:
                line['torch_eager_b200'] = torch_eager_on_gpu(args.seq_len)
            except Exception as exc:                          # informational only
                line['torch_eager_b200'] = {'error': str(exc)}
    print(json.dumps(line))
    if wsize > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_native(args)


if __name__ == '__main__':
    main()
