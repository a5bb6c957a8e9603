"""One rank's shard of the strong-scaled sweep, alone on one GPU: device time per epoch for a few library knobs.
usage: python profiles/shard_ab.py [world=8] [rank=0] [epochs=400] [single]   (env knobs are set per variant inside;
"single": one variant, the knobs of the calling environment -- for knobs that are read once per process)"""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'nerf-attention_b200')); sys.path.insert(0, str(ROOT))
import torch
import bench
from nerf_attention import batched

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 400
specs = bench.sweep_specs(rank, world, 'strong', 2048)
jobs = bench.build_jobs(specs, 2048, pin=False)
initial = []
for j in jobs:
    flat = torch.empty(j.model.count_parameters(), dtype=torch.float32)
    batched.pack_model(j.model, flat)
    initial.append(flat)
single = len(sys.argv) > 4 and sys.argv[4] == 'single'
variants = [('default', {}), ('no_pack', {'NERFATTN_NO_PACK': '1'}), ('no_resident', {'NERFATTN_NO_RESIDENT': '1'})]
if single:
    variants = [(' '.join(f'{k}={v}' for k, v in sorted(os.environ.items()) if k.startswith('NERFATTN_')) or 'default', None)]
for name, env in variants * 2:
    if env is not None:
        for k in ('NERFATTN_NO_RESIDENT', 'NERFATTN_RESIDENT_MAX_H', 'NERFATTN_NO_PACK'):
            os.environ.pop(k, None)
        os.environ.update(env)
    for j, flat in zip(jobs, initial):
        batched.adopt_packed(j.model, flat)
    b = batched.FitBatch(jobs, epochs=epochs, device='cuda', precision='bf16', keep_initial=True)
    best = None
    for _ in range(3):
        b.reset(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.launch(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); best = ms if best is None else min(best, ms)
    res = b.collect()
    print(json.dumps({'variant': name, 'fits': len(jobs), 'world': world, 'rank': rank, 'ms_per_epoch': best / epochs,
                      'fit_epochs_per_sec_x_world': world * len(jobs) * epochs / (best / 1e3) * (280 / (world * len(jobs))),
                      'cos_mean': sum(r.final_cosine_mean for r in res) / len(res)}), flush=True)
    del b
    torch.cuda.empty_cache()
