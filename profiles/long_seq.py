"""BASELINE config 5 (shape): `medium` fits at long sequence lengths, fits/epochs bounded so the run
takes seconds.  Prints fit-epochs/s and TFLOP/s per sequence length for both precisions."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'nerf-attention_b200'))
import torch
import nerf_attention as na
from nerf_attention import batched
from nerf_attention.extract import synthetic_head

tag = sys.argv[1] if len(sys.argv) > 1 else 'r01'
cfg = next(c for c in na.CONFIGS_FULL if c.name == 'medium')
rows = []
for n, nfits, epochs in ((4096, 128, 60), (8192, 64, 60), (16384, 32, 60), (32768, 16, 60)):
    tensors = [synthetic_head(l, 0, n, 32, 8, 128) for l in range(2)]
    flat = [t for kv in tensors for t in kv]
    for prec in ('bf16', 'fp32'):
        torch.manual_seed(0)
        jobs = [na.FitJob(flat[i % len(flat)], cfg) for i in range(nfits)]
        b = batched.FitBatch(jobs, epochs=epochs, device='cuda', precision=prec, keep_initial=True)
        b.launch(); torch.cuda.synchronize()
        b.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); b.launch(); e1.record(); torch.cuda.synchronize()
        sec = e0.elapsed_time(e1) / 1e3
        res = b.collect()
        rows.append({'seq_len': n, 'fits': nfits, 'epochs': epochs, 'precision': prec,
                     'fit_epochs_per_sec': nfits * epochs / sec,
                     'tflops': sum(b.flops) * epochs / sec / 1e12, 'cos_mean': float(sum(r.final_cosine_mean for r in res) / nfits)})
        print(rows[-1])
        del b
        torch.cuda.empty_cache()
(Path(__file__).parent / f'long_seq_{tag}.json').write_text(json.dumps(rows, indent=1))
