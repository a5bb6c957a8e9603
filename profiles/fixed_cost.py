"""Fixed cost of one batched call on a shard of the strong-scaled sweep: device time of reset + launch for several epoch
counts (slope = per-epoch cost, intercept = set-up + final evaluation + graph capture), host time of launch(), and
back-to-back steps as bench.py issues them.
usage: python profiles/fixed_cost.py [world=8] [rank=0]"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'nerf-attention_b200')); sys.path.insert(0, str(ROOT))
import torch
import bench
from nerf_attention import batched

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
specs = bench.sweep_specs(rank, world, 'strong', 2048)
jobs = bench.build_jobs(specs, 2048, pin=False)
initial = []
for j in jobs:
    flat = torch.empty(j.model.count_parameters(), dtype=torch.float32)
    batched.pack_model(j.model, flat)
    initial.append(flat)
rows = []
for epochs in (200, 400, 1000, 2000):
    for j, flat in zip(jobs, initial):
        batched.adopt_packed(j.model, flat)
    b = batched.FitBatch(jobs, epochs=epochs, device='cuda', precision='bf16', keep_initial=True)
    best = None; host = None
    for _ in range(3):
        b.reset(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter(); b.launch(); t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if best is None or ms < best:
            best, host = ms, 1e3 * (t1 - t0)
    # back to back, as bench.py's timed loop
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        b.reset(); b.launch()
    e1.record(); torch.cuda.synchronize()
    rows.append({'fits': len(jobs), 'epochs': epochs, 'device_ms': best, 'launch_host_ms': host, 'back_to_back_ms_per_step': e0.elapsed_time(e1) / 3})
    print(json.dumps(rows[-1]), flush=True)
    b.collect(); del b; torch.cuda.empty_cache()
a, c = rows[1], rows[3]
slope = (c['device_ms'] - a['device_ms']) / (c['epochs'] - a['epochs'])
print(json.dumps({'ms_per_epoch': slope, 'fixed_ms': a['device_ms'] - slope * a['epochs']}))
