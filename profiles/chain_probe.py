"""Chain kernels / dW + Adam kernels of the 280-fit sweep alone (profiling build, NERFATTN_PHASE), per architecture
subset and per set of debug / experiment knobs: ms per epoch, live CUDA events.
usage: python profiles/chain_probe.py <lib.so> <archs: all | medium,hifreq,..> [K=V,K=V ...]   one line per knob set"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'nerf-attention_b200')); sys.path.insert(0, str(ROOT))
import torch
import bench
import nerf_attention as na
from nerf_attention import batched, _native

lib_path = sys.argv[1]
archs = None if sys.argv[2] == 'all' else set(sys.argv[2].split(','))
knobsets = sys.argv[3:] or ['']
os.environ['NERFATTN_PROF_LIB'] = str(Path(lib_path).resolve())
specs = bench.sweep_specs(0, 1, 'strong', 2048)
jobs = bench.build_jobs(specs, 2048, pin=False)
if archs:
    jobs = [j for j in jobs if j.config.name in archs]
initial = []
for j in jobs:
    flat = torch.empty(j.model.count_parameters(), dtype=torch.float32)
    batched.pack_model(j.model, flat)
    initial.append(flat)
pe = 60
for ks in knobsets:
    kv = dict(x.split('=') for x in ks.split(',') if x)
    os.environ.update(kv)
    for j, flat in zip(jobs, initial):
        batched.adopt_packed(j.model, flat)
    b = batched.FitBatch(jobs, epochs=pe, device='cuda', precision='bf16', keep_initial=True, lib=_native.prof_lib())

    def phase_ms(mask):
        os.environ['NERFATTN_PHASE'] = str(mask)
        best = None
        for _ in range(3):
            b.reset(); torch.cuda.synchronize()
            p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            p0.record(); b.launch(); p1.record(); torch.cuda.synchronize()
            ms = p0.elapsed_time(p1); best = ms if best is None else min(best, ms)
        os.environ.pop('NERFATTN_PHASE', None)
        return best
    base = phase_ms(8)
    rec = {'knobs': ks or 'default', 'fits': len(jobs), 'chain_ms_per_epoch': (phase_ms(1) - base) / pe,
           'dw_adam_ms_per_epoch': (phase_ms(2) - base) / pe, 'all_ms_per_epoch': (phase_ms(7) - base) / pe}
    print(json.dumps(rec), flush=True)
    b.collect(); del b; torch.cuda.empty_cache()
    for k in kv:
        os.environ.pop(k, None)
